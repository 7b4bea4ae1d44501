#!/bin/bash
# round 2, GPU call AA: late PDL trigger in the tcgen05 kernels (variant lib), SM split for the SA1-sized wgrads, BN-reduce
# blocks per SM under overlap
mkdir -p gpurun_out
T0=$SECONDS
L=open3d-pointnet2-semantic3d_b200/lib
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/aa_bench_$tag.json 2> gpurun_out/aa_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0)) $(tail -c 300 gpurun_out/aa_bench_$tag.err | tr '\n' ' ')"; }
run base A=1
run lt PN2_LIB=$L/libpn2_b200_lt.so
run base2 A=1
run lt2 PN2_LIB=$L/libpn2_b200_lt.so
run big96 PN2_WGRAD_SMS_BIG=96
run big80 PN2_WGRAD_SMS_BIG=80
run big48 PN2_WGRAD_SMS_BIG=48
run bps3 PN2_BNRED_BPS=3
run bps4 PN2_BNRED_BPS=4
run base3 A=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/aa_bench_*.json")):
    tag = f.split("aa_bench_")[1][:-5]
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-8s %.3f ms/step value %.4g e2e %.4g graph %s loss %.4f" % (
            tag, d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["cuda_graph"], d["e2e"]["last_loss"]))
    except Exception as e:
        print(tag, "parse error", e)
PY
PN2_LIB=$L/libpn2_b200_lt.so timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/aa_suite_lt.log 2>&1; echo "suite(lt) rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/aa_suite_lt.log)"; grep -E "^FAILED|^ERROR" gpurun_out/aa_suite_lt.log | cut -c1-300
