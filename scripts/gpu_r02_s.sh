#!/bin/bash
# round 2, GPU call S: programmatic dependent launch, second pass (the SIMT GEMM in the header had no wait); A/B of
# the trigger placement: none (nt), everywhere but FPS (default build), everywhere (fpst), and PN2_PDL=0
mkdir -p gpurun_out
T0=$SECONDS
L=open3d-pointnet2-semantic3d_b200/lib
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/s_suite.log 2>&1; rc=$?; echo "suite(pdl on) rc=$rc t=$((SECONDS-T0)) $(tail -1 gpurun_out/s_suite.log)"; grep -E "^FAILED|^ERROR" gpurun_out/s_suite.log | cut -c1-300
if [ $rc != 0 ]; then
  PN2_PDL=0 timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/s_suite_pdl0.log 2>&1; echo "suite(pdl off) rc=$? t=$((SECONDS-T0)) $(tail -1 gpurun_out/s_suite_pdl0.log)"; grep -E "^FAILED|^ERROR" gpurun_out/s_suite_pdl0.log | cut -c1-300
fi
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/s_bench_$tag.json 2> gpurun_out/s_bench_$tag.err; echo "bench $tag rc=$? t=$((SECONDS-T0))"; }
for rep in a b; do
run pdl0$rep PN2_PDL=0
run dflt$rep PN2_PDL=1
run nt$rep PN2_PDL=1 PN2_LIB=$L/libpn2_b200_nt.so
run fpst$rep PN2_PDL=1 PN2_LIB=$L/libpn2_b200_fpst.so
done
python - <<'PY'
import json
for tag in ("pdl0a", "pdl0b", "dflta", "dfltb", "nta", "ntb", "fpsta", "fpstb"):
    try:
        d = json.loads(open("gpurun_out/s_bench_%s.json" % tag).read().strip().splitlines()[-1])
        bd = d["breakdown_ms_per_step"]
        print("%-6s %.3f ms/step e2e %.4g | eager sum %.3f | clocks %s" % (
            tag, d["ms_per_step"], d["e2e"]["value"], sum(v["ms_per_step"] for v in bd.values()), d["clocks"].get("sm_mhz")))
    except Exception as e:
        print(tag, "parse error", e)
PY
