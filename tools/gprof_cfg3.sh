set -x
D=$(mktemp -d); python tools/make_newmsm_case.py --out $D --level 6 --D 40 > /dev/null; echo "--numthreads=16" >> $D/conf_MSMAllStrain; mkdir -p $D/out
cd $D && OMP_NUM_THREADS=16 /root/repo/integration/_build/newmsm_gpu_pg --inmesh=sphere.asc --refmesh=sphere.asc --indata=indata.txt --refdata=refdata.txt --conf=conf_MSMAllStrain --out=out/ -f ASCII > /dev/null 2>&1
gprof -b -p /root/repo/integration/_build/newmsm_gpu_pg gmon.out 2>/dev/null | head -45 > /root/repo/gpurun_out/gprof_cfg3_flat.txt
gprof -b -q /root/repo/integration/_build/newmsm_gpu_pg gmon.out 2>/dev/null | grep -E "^\[[0-9]+\]" | head -60 > /root/repo/gpurun_out/gprof_cfg3_incl.txt
