#!/bin/bash
# per-kernel times of the top phase for (order, aggregation) variants
for v in ${VARIANTS:-"0 0" "1 0" "1 1"}; do set -- $v
MSMGPU_BUILD_TOP_ORDER=$1 MSMGPU_BUILD_TOP_AGG=$2 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'^k_top_(count|fill|sort)' -c 6 --csv --log-file gpurun_out/${TAG}_$1_$2.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-parity --no-cpu-baseline --no-adapter-e2e --no-gmsm --no-unary --no-newmsm > /dev/null 2>&1
python - <<EOF
import csv
lines=[l for l in open("gpurun_out/${TAG}_$1_$2.csv") if not l.startswith("==")]
out={}
for x in csv.DictReader(lines):
    out.setdefault(x["Kernel Name"].split("(")[0],{})[x["Metric Name"]]=x["Metric Value"]
print("order=$1 agg=$2", out)
EOF
done
