import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        r = d['roofline']
        print('ms/step %.2f  seg/s %.1f  e2e %.1f  gateTF %.0f (%.0f%%)  classes %s  synth %.0f kHz  launches %d' % (
            d['ms_per_step'], d['value'], d['e2e']['value'], r['achieved'], 100 * r['frac'],
            r['kernel_classes_ms_per_step'], d['synth']['value'] if d.get('synth') else -1, d['gpu_launches']))
