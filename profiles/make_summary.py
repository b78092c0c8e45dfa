"""Rebuilds profiles/r01_final_summary.md and profiles/r01_traffic.json from the artifacts next to it:
r01_final_launches_ncu.csv (ncu launch list of bench.py --profile_mode), r01_final_ncu_full.jsonl (metrics extracted
from the ncu --set full captures: `ncu --set full --clock-control none -k regex:<kernel> -c 2 -o rep python
profiles/tools/prof_step.py; ncu -i rep.ncu-rep --page raw --csv | python profiles/tools/ncu_extract.py`), r01_bench_final.json (bench.py)."""
import collections, csv, json, os, re
H = os.path.dirname(os.path.abspath(__file__))
P = lambda n: os.path.join(H, n)

lines = [l for l in open(P('r01_final_launches_ncu.csv')) if not l.startswith('==')]
r = list(csv.DictReader(lines))
names = [re.sub(r'\(.*', '', x['Kernel Name']) for x in r]
vals = [float(x['Metric Value']) / 1e3 for x in r]
pos = [i for i, n in enumerate(names) if 'loss_pl' in n]
per = pos[-1] - pos[-2]
step = list(zip(names[len(names) - per:], vals[len(names) - per:]))
agg = collections.OrderedDict()
for k, v in step:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
tab = ["| share | total us | launches | avg us | kernel |", "|---|---|---|---|---|"]
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    tab.append("| %.1f%% | %.0f | %d | %.1f | `%s` |" % (100 * v / tot, v, c, v / c, k[:80]))
ours = sum(1 for k, _ in step if 'intel::' in k or 'umma::' in k)
tab.append("")
tab.append("launches in the step: %d (of which ours: %d), sum of kernel times %.0f us" % (len(step), ours, tot))

recs = [json.loads(l) for l in open(P('r01_final_ncu_full.jsonl'))]
f = lambda rr, k: (rr.get(k, '') or '-').split()[0]
seen, rows, traffic = set(), [], {}
name_map = {'trunk_bwd_kernel': 'trunk_bwd', 'trunk_fwd_kernel': 'trunk_fwd', 'gru_seq_fwd_kernel': 'gru_seq_fwd',
            'gru_seq_bwd_kernel': 'gru_seq_bwd', 'dense_rows_fwd_kernel': 'dense_rows_fwd', 'gemm_umma_kernel': 'gemm_fwd'}
for rr in recs:
    k = rr['kernel'].split('(')[0].replace('void ', '')
    if k in seen:
        continue
    seen.add(k)
    dr, dw = float(f(rr, 'dram__bytes_read.sum')), float(f(rr, 'dram__bytes_write.sum'))
    tens = rr.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
                  f(rr, 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active'))
    rows.append("| `%s` | %s | %s | %.1f %% | %.1f %% | %.1f %% | %.0f / %.0f MB | %s |" % (
        k[:48], f(rr, 'gpu__time_duration.sum'), f(rr, 'launch__registers_per_thread'),
        float(f(rr, 'sm__warps_active.avg.pct_of_peak_sustained_active')),
        float(f(rr, 'smsp__issue_active.avg.pct_of_peak_sustained_active')), float(tens or 0), dr, dw,
        ", ".join("%s %.0f" % (a, b) for a, b in list(rr['stalls_pct'].items())[:4])))
    for kk, vv in name_map.items():
        if kk in rr['kernel'] and vv not in traffic:
            traffic[vv] = {"dram_bytes_per_launch": int((dr + dw) * 1e6), "kernel": rr['kernel'],
                           "time_us_under_ncu": float(f(rr, 'gpu__time_duration.sum')),
                           "source": "profiles/r01_final_ncu_full.jsonl (ncu --set full --clock-control none, one train step)"}
json.dump(traffic, open(P('r01_traffic.json'), 'w'), indent=1)

d = json.loads(open(P('r01_bench_final.json')).read().strip().splitlines()[-1])
pk = d['roofline']['per_kernel']
live = ["| kernel | share | ms/step | GB/s (algorithmic) | TFLOP/s (3xTF32 issued) | launches/step |", "|---|---|---|---|---|---|"]
for k, v in list(pk.items())[:16]:
    live.append("| %s | %.1f%% | %.3f | %.0f | %s | %.0f |" % (k, 100 * v['share'], v['share'] * d['ms_per_step'], v['gbs'],
                ("%.0f" % v['tflops_3xtf32']) if v.get('tflops_3xtf32') else '-', v['launches_per_step']))
e = d['e2e']
md = f"""# Round 1, final state - B200, IntEL-PL flags (`script/IntEL.sh`), B=4096 sessions, L=50, K=4, I=1071

`bench.py` (r01_bench_final.json): value = {d['value']:.0f} sessions/s ({d['ms_per_step']:.2f} ms/step, inputs resident in HBM);
e2e = {e['value']:.0f} sessions/s from pinned host batches in the reference's dense float64 layout
({e['input_layout']}; {e['h2d_bytes_per_step'] / 1e6:.0f} MB over PCIe per step; modes: dense copy
{e['modes']['dense_copy']['value']:.0f}, host packed {e['modes']['host_packed']['value']:.0f} sessions/s);
e2e_compact = {d['e2e_compact']['value']:.0f} sessions/s; eval = {d['eval_sessions_per_s']:.0f} sessions/s;
CPU port of the reference = {d['cpu_baseline']['value']:.0f} sessions/s on {d['cpu_baseline']['cores']} cores
(`--impl reference`: r01_bench_reference_arm.json); clocks {d['clocks']}; {d['gpu_launches'] // d['steps']} of our kernels per step.
Weak scaling (r01_bench_final_{{2,4,8}}gpu.json): 1.48 M / 2.93 M / 5.83 M sessions/s on 2 / 4 / 8 GPUs (5.55 / 5.59 / 5.62 ms per step, 95 % at 8).

The first-path summary of this round is kept in r01_summary.md (16.4 ms/step); milestones in between are in DESIGN.md section 10.

## ncu launch list of one train step
`ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv python bench.py --steps 2 --warmup 1 --profile_mode`
(raw list: r01_final_launches_ncu.csv; the last of the three steps is shown; cold-cache, serialised: compare SHARES with the live table)

{chr(10).join(tab)}

## live CUDA-event shares from bench.py (same kernels, warm, events around every launch)

{chr(10).join(live)}

The two tables agree on the shares and on the total (sum of ncu kernel times vs the live step time).

## `ncu --set full --clock-control none` of the top kernels (one train step; extracted metrics in r01_final_ncu_full.jsonl)

| kernel | time us | regs | warps active | issue active | tensor pipe active | DRAM read / write | top stall reasons (% of samples) |
|---|---|---|---|---|---|---|---|
{chr(10).join(rows)}

Reading:
* `trunk_bwd` / `trunk_fwd` / `gru_seq_*` run on the legacy HMMA pipe (mma.sync TF32) at 31-49 % pipe activity with one
  CTA per SM (8-16 warps): what remains are fixed-latency waits between dependent MMAs and the issue slots spent on the
  hi/lo splits.  DRAM traffic of trunk_bwd (420 MB) equals its algorithmic bytes (the saved activations): no wasted
  re-reads; it moves them at 0.6 TB/s because the tensor pipe, not HBM, paces the kernel.  Two restructurings were
  measured and dropped: issuing the three MMAs of a product in separate passes (+23 % time: register spills) and
  hoisting all loads of a phase (no change: ptxas already overlaps them).
* `gemm_umma_kernel` (tcgen05): the UTCHMMA pipe is ~12 % active; the kernel is paced by the cp.async -> split -> plane
  conversion loop (issue active 52 %, short-scoreboard stalls on shared memory) - the next thing to restructure
  (persistent CTAs with the epilogue of one tile overlapping the main loop of the next).  Measured and dropped:
  64-wide tiles for two CTAs per SM (no gain), an 8-deep cp.async ring (slower: the bottleneck is not memory-level
  parallelism), a dedicated MMA-issuer warp with mbarrier hand-offs (long-K shapes 20-50 % slower: the chunk drain
  serialises harder), pointer-advancing address arithmetic with a full-tile fast path (slower despite fewer
  instructions).  ncu source-level sampling shows no hot spot: stall samples are spread over the whole k-tile body.
* `dense_rows_fwd` streams the live rows of the float64 [B,H,I] history intents (the padding rows behind history_len
  are skipped, one warp per row so that the block scheduler balances live and padding rows): ~370 MB of DRAM reads per
  launch instead of 707 MB; when every row is live it runs at 5.9 TB/s = 0.90 of the measured HBM peak.

r01_traffic.json holds the per-launch DRAM bytes of these captures; `bench.py` copies the entry of the dominant kernel
into `roofline.traffic`.
"""
open(P('r01_final_summary.md'), 'w').write(md)
print(md[:1500])
