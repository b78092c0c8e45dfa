"""Rebuilds profiles/r02_summary.md and profiles/r02_traffic.json from the artifacts next to it:
r02_launches_ncu.csv (ncu launch list of `bench.py --steps 2 --warmup 1 --profile_mode`), r02_ncu_*.jsonl (metrics extracted
from `ncu --set full --clock-control none --import-source on` captures with profiles/tools/ncu_extract.py), r02_bench_*.json
(bench.py lines from the B200 box), r02_probe.log (tests/hw/umma_probe.cu), r02_sass_summary.txt."""
import collections, csv, json, os, re
H = os.path.dirname(os.path.abspath(__file__))
P = lambda n: os.path.join(H, n)


def last_json(name):
    return json.loads(open(P(name)).read().strip().splitlines()[-1])


lines = [l for l in open(P('r02_launches_ncu.csv')) if not l.startswith('==')]
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
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
    tab.append("| %.1f%% | %.0f | %d | %.1f | `%s` |" % (100 * v / tot, v, c, v / c, k[:80]))
ours = sum(1 for k, _ in step if 'intel::' in k or 'umma::' in k)
tab.append("")
tab.append("launches in the step: %d (of which ours: %d), sum of kernel times %.0f us" % (len(step), ours, tot))

f = lambda rr, k: (rr.get(k, '') or '-').split()[0]
rows, traffic, seen = [], {}, set()
name_map = {'trunk_bwd_kernel': 'trunk_bwd', 'trunk_tc_fwd_kernel': 'trunk_fwd', 'gru_tc_fwd_kernel': 'gru_seq_fwd',
            'gru_seq_bwd_kernel': 'gru_seq_bwd', 'dense_rows_fwd_kernel': 'dense_rows_fwd', 'ndcg_kernel': 'ndcg',
            'batch_build_kernel': 'batch_build', 'scatter_add_kernel': 'scatter_add', 'reduce_partials_kernel': 'ndcg_reduce',
            'loss_pl_kernel': 'loss_pl', 'intent_loss_kernel': 'intent_loss', 'gemm_wgrad_tc_kernel': 'gemm_wgrad',
            'gemm_rows_tc_kernel': 'gemm_rows', 'cross_pool_bwd_kernel': 'cross_pool_bwd', 'gemm_umma_kernel': 'gemm_umma'}
for src, note in (('r02_ncu_train.jsonl', 'train step (activations saved)'),
                  ('r02_ncu_eval.jsonl', 'eval step (inference mode: nothing saved) + one device-built batch'),
                  ('r02_ncu_late.jsonl', 'train step, late (the small kernels rewritten after the main capture; these rows replace the older ones above)')):
    if not os.path.exists(P(src)):
        continue
    for line in open(P(src)):
        rr = json.loads(line)
        k = rr['kernel'].split('(')[0].replace('void ', '')
        key = (k, src)
        if key in seen:
            continue
        seen.add(key)
        def mb(key):                            # ncu prints "12.3 Mbyte" / "721.4 Kbyte" / "1.2 Gbyte" / "512 byte"
            t = (rr.get(key, '') or '0 byte').split()
            unit = (t[1] if len(t) > 1 else 'byte').lower()
            return float(t[0]) * {'byte': 1e-6, 'kbyte': 1e-3, 'mbyte': 1.0, 'gbyte': 1e3}.get(unit, 1.0)
        dr, dw = mb('dram__bytes_read.sum'), mb('dram__bytes_write.sum')
        tens = rr.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
                      f(rr, 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active'))
        try:
            tens = float(tens)
        except Exception:
            tens = 0.0
        rows.append("| `%s` | %s | %s | %s | %.1f %% | %.1f %% | %.1f / %.1f MB | %s |" % (
            k[:44], note.split(' (')[0], f(rr, 'gpu__time_duration.sum'), f(rr, 'launch__registers_per_thread'),
            float(f(rr, 'smsp__issue_active.avg.pct_of_peak_sustained_active')), tens, dr, dw,
            ", ".join("%s %.0f" % (a, b) for a, b in list(rr['stalls_pct'].items())[:4])))
        for kk, vv in name_map.items():
            if src == 'r02_ncu_eval.jsonl':
                vv = vv + '@eval'               # the same kernel in inference mode (bench.py: roofline_eval)
            if kk in rr['kernel'] and vv not in traffic:
                traffic[vv] = {"dram_bytes_per_launch": int((dr + dw) * 1e6), "kernel": rr['kernel'],
                               "time_us_under_ncu": float(f(rr, 'gpu__time_duration.sum')),
                               "source": "profiles/%s (ncu --set full --clock-control none; %s)" % (src, note)}
json.dump(traffic, open(P('r02_traffic.json'), 'w'), indent=1)

d = last_json('r02_bench_c2.json')
pk = d['roofline']['per_kernel']
live = ["| kernel | share | ms/step | GB/s (algorithmic) | TFLOP/s (3xTF32 issued) | launches/step |", "|---|---|---|---|---|---|"]
for k, v in list(pk.items())[:18]:
    live.append("| %s | %.1f%% | %.3f | %.0f | %s | %.0f |" % (k, 100 * v['share'], v['ms_per_step'], v['gbs'],
                ("%.0f" % v['tflops_3xtf32']) if v.get('tflops_3xtf32') else '-', v['launches_per_step']))
cfg_rows = ["| config | step | sessions/s (`value`) | ms/step | e2e sessions/s | CPU reference (16 cores) |", "|---|---|---|---|---|---|"]
for c in ('c2', 'c3', 'c4', 'c5'):
    if os.path.exists(P('r02_bench_%s.json' % c)):
        x = last_json('r02_bench_%s.json' % c)
        cb = x.get('cpu_baseline') or {}
        cfg_rows.append("| %s | %s | %.0f | %.3f | %.0f | %s |" % (c, x['config']['step'], x['value'], x['ms_per_step'], x['e2e']['value'],
                        ("%.0f (%s)" % (cb['value'], cb['kind'])) if cb else '-'))
scale = ["| GPUs | config | sessions/s | ms/step | per-GPU efficiency | e2e sessions/s | e2e per-GPU efficiency |", "|---|---|---|---|---|---|---|"]
for c, base in (('c2', 'r02_bench_c2.json'), ('c4', 'r02_bench_c4.json')):
    if not os.path.exists(P(base)):
        continue
    # the multi-GPU lines are compared with the single-GPU line measured in the same session (later kernel work moved the
    # single-GPU line on; the base of the scaling table stays the one the multi-GPU runs were taken with)
    sb = base.replace('.json', '_scaling_base.json')
    if os.path.exists(P(sb)):
        base = sb
    b1 = last_json(base)
    for n, fn in ((1, base), (2, 'r02_bench_%s_2gpu.json' % c if c != 'c2' else 'r02_bench_2gpu.json'),
                  (4, 'r02_bench_%s_4gpu.json' % c if c != 'c2' else 'r02_bench_4gpu.json'),
                  (8, 'r02_bench_%s_8gpu.json' % c if c != 'c2' else 'r02_bench_8gpu.json')):
        if os.path.exists(P(fn)):
            x = last_json(fn)
            scale.append("| %d | %s | %.0f | %.3f | %.3f | %.0f | %.3f |" % (n, c, x['value'], x['ms_per_step'], x['value'] / (n * b1['value']),
                         x['e2e']['value'], x['e2e']['value'] / (n * b1['e2e']['value'])))
ev = d.get('roofline_eval', {})
c4f = last_json('r02_bench_c4_fused_bert.json')
md = f"""# Round 2 - B200, BASELINE.json configs c2..c5 (bench.py --config), IntEL-PL flags for c2 / c3

`bench.py` c2 (r02_bench_c2.json): value = {d['value']:.0f} sessions/s ({d['ms_per_step']:.3f} ms/step of 4096 sessions, inputs resident in HBM,
dense reference layout); e2e = {d['e2e']['value']:.0f} sessions/s (host session indices -> H2D -> `intel_batch_build` from the device-resident
corpus -> step -> loss D2H; median of 3 passes {['%.3f' % p for p in d['e2e']['passes_ms']]} ms); eval = {d.get('eval_sessions_per_s', 0):.0f} sessions/s;
CPU reference = {d.get('cpu_baseline', {}).get('value', 0):.0f} sessions/s ({d.get('cpu_baseline', {}).get('kind')}, {d.get('cpu_baseline', {}).get('cores')} cores);
clocks {d['clocks']}; {d['gpu_launches'] / d['steps']:.0f} kernel launches per step.

## every BASELINE config, 1 GPU
{chr(10).join(cfg_rows)}

(all four lines above are from the final kernels of the round.  The multi-GPU lines below are older: the c2 ones were taken
at a 4.30 ms step, the c4 ones before the fused BERT4Rec encoder and the second warpgroup of the 200-slot stack kernel, which
took the single-GPU c4 line from 1.27 M to {c4f['value']:.0f} sessions/s that day - r02_bench_c4_fused_bert.json,
r02_ncu_bert_fused.jsonl; efficiencies are computed against the single-GPU line of the same session)

## weak scaling (one rank per GPU, NCCL gradient all-reduce for the train step; none for eval; all lines of one config from one session: *_scaling_base.json is the single-GPU line of that session)
{chr(10).join(scale)}

## ncu launch list of one train step
`ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv python bench.py --steps 2 --warmup 1 --profile_mode`
(raw list: r02_launches_ncu.csv; the last step is shown; cold-cache, serialised: compare SHARES with the live table)

{chr(10).join(tab)}

## live CUDA-event shares from bench.py (same kernels, warm, events around every launch)

{chr(10).join(live)}

## eval step (forward under no_grad + intel_ndcg_topk): {ev.get('ms_per_step', 0):.3f} ms per 4096 sessions
dominant kernel: {json.dumps(ev.get('dominant', {}))[:400]}
ndcg kernel: {json.dumps(ev.get('ndcg_kernel', {}))[:400]}

## `ncu --set full --clock-control none --import-source on` (extracted metrics: r02_ncu_*.jsonl)

| kernel | captured in | time us | regs | issue active | tensor pipe active | DRAM read / write | top stall reasons (% of samples) |
|---|---|---|---|---|---|---|---|
{chr(10).join(rows)}

r02_traffic.json holds the per-launch DRAM bytes of these captures; `bench.py` copies the entry of the dominant kernel into
`roofline.traffic`.  r02_sass_summary.txt: per-kernel counts of UTCHMMA / LDTM / STTM / UTCBAR / HMMA in libintel_b200.so.
r02_probe.log: the tcgen05 forms checked bit-exact on the hardware before the kernels were written (tests/hw/umma_probe.cu; 13 and
14 are layouts the hardware rejects, kept as negative results; 15-17: the tensor cores ignore the low 13 operand bits).
r02_trunk_bwd_lines.txt / r02_gru_tc_fwd_lines.txt / r02_gemm_umma_lines.txt: instructions and stall samples per source
statement (profiles/tools/sass_lines.py).  r02_shapes.log: CUDA-event time of every kernel / GEMM shape of one train step.
r02_bench_8gpu_overlap.json / r02_bench_8gpu_no_overlap.json: the gradient exchange hidden behind the backward pass, A/B.
"""
open(P('r02_summary.md'), 'w').write(md)
print(md[:3000])
