for g in 0 1; do for r in 0 1; do python scripts/probe_small.py --graph $g --reorder $r; done; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_small_launches.csv python scripts/probe_small.py --graph 0 --steps 20 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/r2_small_launches.csv').read().splitlines() if l.startswith('"')]
rd=list(csv.reader(rows)); h=rd[0]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
import collections
tot=collections.Counter(); cnt=collections.Counter()
data=rd[1:]
last=data[-13*10:]  # last 10 steps
for r in last:
    n=r[kn].split('(')[0]; v=float(r[mv].replace(',','')); u=r[mu]
    v*= {'ns':1e-3,'us':1,'ms':1e3,'usecond':1,'nsecond':1e-3,'msecond':1e3}.get(u,1)
    tot[n]+=v; cnt[n]+=1
for k,v in tot.most_common(): print(k, cnt[k], round(v/10,2),'us per step')
print('sum per step', round(sum(tot.values())/10,1),'us')
PY
