#!/usr/bin/env python
"""Turn gpurun_out/{launches_r1.csv, prof_*.ncu-rep, bench_full.json, bench_ref.json} into the tracked profiles/ files.
usage: python tools/make_profile_summary.py gpurun_out/prof_r1g.ncu-rep"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.per_cycle_active',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg.per_second']
out = []
for r in rows[2:]:
    d = {w: (r[idx[w]] + ' ' + units[idx[w]]).strip() for w in want if w in idx}
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and 'per_issue_active' in h:
            d[h.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')] = r[i]
    out.append(d)
P = os.path.join(ROOT, "profiles")
json.dump(out, open(os.path.join(P, "r1_ncu_drone_step.json"), "w"), indent=1)
for src, dst in (("launches_r1.csv", "r1_launches.csv"), ("bench_full.json", "r1_bench_n1.json"), ("bench_ref.json", "r1_bench_reference_arm.json")):
    s = os.path.join(ROOT, "gpurun_out", src)
    if os.path.isfile(s):
        open(os.path.join(P, dst), "w").write(open(s).read())
lrows = [r for r in csv.reader(open(os.path.join(P, "r1_launches.csv"))) if len(r) > 5 and r[0].isdigit()]
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in lrows:
    name = r[4].split('(')[0][:80]; tot[name] += float(r[-1]) / 1e3; cnt[name] += 1
allt = sum(tot.values())
lines = ["| kernel | launches | total us | share of all launches |", "|---|---|---|---|"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    lines.append(f"| `{k}` | {cnt[k]} | {v:.1f} | {v / allt:.1%} |")
step = [float(r[-1]) / 1e3 for r in lrows if 'drone_step_tma' in r[4]]
k8 = out[0]
g = lambda d, k: d.get(k, '')
bench = json.load(open(os.path.join(P, "r1_bench_n1.json")))
md = f"""# Round 1 ncu evidence (B200, sm_100a) -- `python bench.py --steps 3 --warmup 3 --profile`

Commands (B200_PROFILING.md recipe; the plain run exited 0 first, in the same gpurun call):
```
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --profile
ncu --set full --clock-control none --import-source on -k regex:drone_step -s 6 -c 4 -o gpurun_out/prof_r1g python bench.py --steps 3 --warmup 3 --profile
```
Files here: `r1_launches.csv` (every launch of the profile run), `r1_ncu_drone_step.json` (raw-page metrics of the
captured `drone_step_tma_kernel` launches, all K = 8), `r1_bench_n1.json` / `r1_bench_reference_arm.json` (bench lines of
the same build taken WITHOUT a profiler), `r1_fp32_pipe_microbench.txt` (tools/fma_microbench.cu),
`r1_warp_timeline_static.txt` (per-warp start/end with the static chunk split, the motivation for pulled chunks),
`r1_chain_timeline.txt` (per-warp start/end of three chained launches), `r1_batch_size_sweep.txt` (step time vs batch size:
steady-state cost and fixed cost), `r1_chase_bench.json` / `r1_chase_launches.csv` / `r1_ncu_camera_splat.json` (chase pipeline).

## Launch list (ncu serialises and flushes caches: compare shares, not absolutes)
Inside a timed step there is exactly ONE launch, `drone_step_tma_kernel` (share of the step: 100 %; `gpu_launches` = steps).
In the timed rotation the launches are CHAINED (DESIGN.md section 4.1): launch i+1 starts on the SMs launch i has left, which a
serialising profiler cannot show -- `r1_chain_timeline.txt` (tools/tune_chain_trace.py, %globaltimer per warp) does.
The rest of the profile run is set-up (synthetic init) and the L2-flush kernels of the cross-check loop.

{chr(10).join(lines)}

Per-launch durations of `drone_step_tma_kernel` in launch order (us): {', '.join(f'{x:.1f}' for x in step)}
(order, each 3 warm-up + 3 timed: K=8 rotation chained on 2 of 4 CTA slots [grid 296: under ncu's serialisation a half-size
grid simply takes longer -- side by side they overlap], K=8 rotation chained full grid, K=8 rotation unchained, K=8 flushed,
then the same three rotations and the flushed loop at K=1; the full capture below is of the full-grid launches).

## `fpv::drone_step_tma_kernel<F2, ANG=4, 128, 4, 2>` -- K = 8, 1,048,576 envs

| metric | value |
|---|---|
| duration under ncu | {g(k8, 'gpu__time_duration.sum')} (CUDA events in bench.py, no profiler: {bench['ms_per_step'] * 1e3:.1f} us back to back, {bench['ms_per_step_flushed'] * 1e3:.1f} us with explicit flush) |
| grid x block | {g(k8, 'launch__grid_size')} x {g(k8, 'launch__block_size')} (= 148 SMs x 4 resident CTAs, persistent) |
| registers / thread | {g(k8, 'launch__registers_per_thread')} |
| dynamic smem / CTA | {g(k8, 'launch__shared_mem_per_block_dynamic')} (occupancy limit: smem {g(k8, 'launch__occupancy_limit_shared_mem')}, regs {g(k8, 'launch__occupancy_limit_registers')}) |
| DRAM read / write | {g(k8, 'dram__bytes_read.sum')} / {g(k8, 'dram__bytes_write.sum')} per launch (algorithmic 152.0 MB; the state stores are still in L2 at kernel end) |
| DRAM throughput | {g(k8, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} of peak |
| warp instructions | {g(k8, 'smsp__inst_executed.sum')} |
| issue slots busy | {g(k8, 'smsp__issue_active.avg.per_cycle_active')} per cycle per scheduler |
| FMA pipe: instructions / cycles active | {g(k8, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active')} / {g(k8, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active')} |
| ALU pipe instructions | {g(k8, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active')} |
| resident warps / SM | {g(k8, 'sm__warps_active.avg.per_cycle_active')} |
| SM clock during capture | {g(k8, 'sm__cycles_elapsed.avg.per_second')} |
| top stalls (per issue) | math_pipe_throttle {g(k8, 'stall_math_pipe_throttle')}, wait {g(k8, 'stall_wait')}, not_selected {g(k8, 'stall_not_selected')}, long_scoreboard {g(k8, 'stall_long_scoreboard')}, dispatch {g(k8, 'stall_dispatch_stall')} |

SASS evidence of the Blackwell-specific paths (`cuobjdump -sass fpyv_b200/libfpyv_b200.so`): `FFMA2/FMUL2/FADD2`
(packed FP32), `UBLKCP.S.G` (TMA bulk copy), `SYNCS.ARRIVE.TRANS64` / `SYNCS.PHASECHK.TRANS64.TRYWAIT` (mbarrier),
`ELECT`, `ACQBULK` / `PREEXIT` (programmatic dependent launch), `REDG.E.ADD.F64` (statistics).

## FP32 pipe micro-benchmark (`tools/fma_microbench.cu`, same box class)
```
{open(os.path.join(P, 'r1_fp32_pipe_microbench.txt')).read().strip()}
```
Reading: packed FFMA2 has the lane throughput of scalar FFMA (no 2x) and costs 2 dispatch cycles -- 3 with three
distinct register-pair operands.  The substep loop (64 FFMA2 + 27 FMUL2 + 6 FADD2 per iteration, ~39 of the FFMA2 with
three register pairs) needs ~233 cycles per warp-iteration of register-file operand bandwidth; see DESIGN.md section 4.
"""
open(os.path.join(P, "r1_ncu_summary.md"), "w").write(md)
print("profiles/ refreshed:", g(k8, 'gpu__time_duration.sum'), g(k8, 'dram__bytes_read.sum'), g(k8, 'dram__bytes_write.sum'))
