#!/bin/bash
# A/B of the octree top phase on the GPU box: tests, value-only bench per knob, ncu launch list of the build kernels
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "octree or nearest or forest" 2>&1 | tail -5
for t in ${TOPS:-0 -1}; do
  MSMGPU_BUILD_TOP=$t MSMGPU_BUILD_TIMING=1 python bench.py --steps 5 --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-adapter-e2e --no-gmsm --no-unary --no-newmsm > gpurun_out/${TAG}_top$t.json 2> gpurun_out/${TAG}_top$t.err
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_top$t.json').read().strip().splitlines()[-1]); print('top', $t, d['ms_per_step'], d['detail']['breakdown_ms_per_step'])"
  tail -2 gpurun_out/${TAG}_top$t.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_(top|chunk|scatter|node|make|scan|save|mesh_tables|init)' -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-parity --no-cpu-baseline --no-adapter-e2e --no-gmsm --no-unary --no-newmsm > gpurun_out/${TAG}_ncu.log 2>&1
python - <<EOF
import csv, collections
lines=[l for l in open("gpurun_out/${TAG}_launches.csv") if not l.startswith("==")]
rows=[(x["Kernel Name"].split("(")[0][:40], float(x["Metric Value"].replace(",",""))/1e3) for x in csv.DictReader(lines)]
# one build = from k_mesh_tables to the next k_mesh_tables
starts=[i for i,r in enumerate(rows) if r[0].startswith("k_mesh_tables")]
seg=rows[starts[-2]:starts[-1]] if len(starts)>=2 else rows
agg=collections.OrderedDict()
for k,v in seg:
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,v) in agg.items(): print(f"{k:42s} x{c:3d} {v:9.1f} us")
print("total", sum(v for _,v in seg))
EOF
