import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']
        print('value %.3f Gcu/s  ms/step %.2f  '%(d['value']/1e9,d['ms_per_step'])+'  '.join('%s %.3f'%(n[2:],v['ms_per_launch']) for n,v in k.items()))
