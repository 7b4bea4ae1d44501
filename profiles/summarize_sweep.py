"""gpurun_out/op_sweep.json (written by profiles/op_sweep.py on the GPU box) -> the tracked table
profiles/op_sweep_<tag>.md.  Run from the repo root: python profiles/summarize_sweep.py r01"""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
d = json.load(open("gpurun_out/op_sweep.json"))
hbm = d["hbm_gbs"]
out = ["# Op-level microbenchmarks on one B200 (`profiles/op_sweep.py`, %s)" % tag, "",
       "CUDA-event times through the public op surface, inputs resident, 256 MB L2 flush before each timed "
       "call, median of up to 5.  HBM peak used for the fractions: %.0f GB/s (MEASURED_PEAKS.json copy "
       "bandwidth).  `ref kernel` = the reference's own tf_ops CUDA kernel compiled unmodified for sm_100a "
       "and timed on the same GPU (oracle/_ref)." % hbm, ""]


def fmt(v, spec="%.3g"):
    return "-" if v is None else spec % v


rows = d["results"]
out += ["## farthest_point_sample (serial-latency bound; streaming model = B*(npoint-1)*N*20 B, SURVEY 8d)", "",
        "| case | B | N | npoint | ms | us/round | streaming-model GB/s | x HBM peak | compulsory GB/s | ref kernel ms | speed-up |",
        "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for r in rows:
    if r.get("op") == "farthest_point_sample":
        out.append("| %s | %d | %d | %d | %.3f | %.3f | %.0f | %.2f | %.2f | %s | %s |" % (
            r["case"], r["B"], r["N"], r["npoint"], r["ms"], r["us_per_round"], r["streaming_model_GBps"],
            r["streaming_model_frac_of_hbm"], r["compulsory_GBps"], fmt(r.get("ref_kernel_ms"), "%.3f"),
            fmt(r.get("speedup_vs_ref_kernel"), "%.1fx")))
out += ["", "## query_ball_point (FP32-issue bound brute force; compulsory bytes B*(12n+12m+4*m*ns+4m))", "",
        "| case | B | n | m | nsample | ms | pair tests/s | frac of FP32 issue (7 instr/pair) | compulsory GB/s | frac of HBM | ref kernel ms | speed-up |",
        "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for r in rows:
    if r.get("op") == "query_ball_point":
        out.append("| %s | %d | %d | %d | %d | %.3f | %.3g | %.2f | %.1f | %.4f | %s | %s |" % (
            r["case"], r["B"], r["n"], r["m"], r["nsample"], r["ms"], r["pair_tests_per_s"],
            r["frac_of_fp32_issue"], r["compulsory_GBps"], r["compulsory_frac_of_hbm"],
            fmt(r.get("ref_kernel_ms"), "%.3f"), fmt(r.get("speedup_vs_ref_kernel"), "%.1fx")))
out += ["", "## three_nn (FP64-ALU bound: exact fp64 distances, 8 DP ops per pair)", "",
        "| case | B | n (queries) | m (known) | ms | pair tests/s | DP op/s | compulsory GB/s | frac of HBM |",
        "|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
for r in rows:
    if r.get("op") == "three_nn":
        out.append("| %s | %d | %d | %d | %.3f | %.3g | %.3g | %.1f | %.4f |" % (
            r["case"], r["B"], r["n"], r["m"], r["ms"], r["pair_tests_per_s"], r["fp64_ops_per_s"],
            r["compulsory_GBps"], r["compulsory_frac_of_hbm"]))
out += ["", "## gathers (HBM bound)", "", "| op | case | ms | GB/s (algorithmic bytes) | frac of HBM | ref kernel ms | speed-up |",
        "|---|---|---:|---:|---:|---:|---:|"]
for r in rows:
    if r.get("op") in ("group_point", "three_interpolate"):
        out.append("| %s | %s | %.3f | %.0f | %.2f | %s | %s |" % (
            r["op"], r["case"], r["ms"], r["GBps"], r["frac_of_hbm"], fmt(r.get("ref_kernel_ms"), "%.3f"),
            fmt(r.get("speedup_vs_ref_kernel"), "%.1fx")))
other = [r for r in rows if r.get("op") in ("prob_sample", "sweep")]
if other:
    out += ["", "## other", ""]
    for r in other:
        out.append("* `%s`" % json.dumps(r))
open("profiles/op_sweep_%s.md" % tag, "w").write("\n".join(out) + "\n")
print("wrote profiles/op_sweep_%s.md" % tag)
