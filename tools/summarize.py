#!/usr/bin/env python
"""Summarise gpurun_out/parity_report.json and ncu launch lists."""
import csv, json, sys
def parity(fn):
    r=json.load(open(fn))
    for x in r:
        bad={k:x[k] for k in ('radii_mismatch','means2D_bit_mismatch','depths_bit_mismatch','conic_bit_mismatch','point_list_mismatch','ranges_mismatch','gaussians_count_mismatch','cov3D_bit_mismatch','img_n_gt_1e-4') if x.get(k)}
        t={k:round(x[k]['ms_median'],3) for k in ('time_ref','time_ours','time_bwd_ref','time_bwd_ours') if k in x}
        g={k:float('%.1e'%v['rel_l2']) for k,v in x.get('grads',{}).items()}
        print(x['variant'],x['tag'],x.get('gaze',''),'N',x.get('num_rendered_ref'),x.get('num_rendered_ours'),'img',x.get('img_max_abs'),'lazy',x.get('lazy_img_max_abs'),(x.get('lazy_stats') or {}).get('blend_consumed'),'BAD' if bad else 'ok',bad,t, 'grad_max_rel', max(g.values()) if g else '', x.get('error',''))
def launches(fn, frames=3):
    with open(fn) as f: lines=[l for l in f if not l.startswith('==')]
    rows=[(x['Kernel Name'], float(x['Metric Value'].replace(',',''))) for x in csv.DictReader(lines)]
    n=len(rows)//frames; last=rows[-n:]; tot=sum(v for _,v in last)
    for k,v in last: print('  %-64s %9.1f us %5.1f%%'%(k[:64], v/1000, 100*v/tot))
    print('  total %.1f us'%(tot/1000))
if __name__=='__main__':
    for a in sys.argv[1:]:
        print('==',a)
        (parity if a.endswith('.json') else launches)(a)
