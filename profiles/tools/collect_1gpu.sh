#!/bin/bash
# Single-GPU evidence run on the B200 box (through gpurun): every BASELINE config of bench.py, the reference arm, the ncu
# launch list of a train step and `ncu --set full` captures of the kernels the summary quotes (metrics extracted on the box:
# the .ncu-rep files are too large to bring back).  Output: gpurun_out/<tag>_*; copy what is to be kept into profiles/.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
for c in c2 c3 c4 c5; do
  python bench.py --config $c > $O/${TAG}_bench_$c.json 2> $O/${TAG}_bench_$c.err
done
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_launches_ncu.csv \
    python bench.py --steps 2 --warmup 1 --profile_mode > $O/${TAG}_launch.log 2>&1
K='trunk_bwd|trunk_tc_fwd|gru_seq_bwd|gru_tc_fwd|gemm_wgrad_tc|gemm_rows_tc|loss_pl|intent_loss|dense_rows_fwd|scatter_add|cross_pool_bwd'
STEPS=2 EVAL=0 timeout 900 ncu --set full --clock-control none -k regex:"$K" -s 30 -c 30 -f -o /tmp/${TAG}_train python profiles/tools/prof_step.py > $O/${TAG}_ncu_train.log 2>&1
ncu -i /tmp/${TAG}_train.ncu-rep --page raw --csv 2>/dev/null | python profiles/tools/ncu_extract.py > $O/${TAG}_ncu_train.jsonl
STEPS=1 EVAL=1 timeout 900 ncu --set full --clock-control none -k regex:'trunk_tc_fwd|gru_tc_fwd|ndcg_kernel|reduce_partials|batch_build|gemm_umma' -s 8 -c 16 -f -o /tmp/${TAG}_eval python profiles/tools/prof_step.py > $O/${TAG}_ncu_eval.log 2>&1
ncu -i /tmp/${TAG}_eval.ncu-rep --page raw --csv 2>/dev/null | python profiles/tools/ncu_extract.py > $O/${TAG}_ncu_eval.jsonl
bash tests/hw/run_probe.sh > $O/${TAG}_probe.log 2>&1
PROBES="12 13 14 15 16 17" bash tests/hw/run_probe.sh >> $O/${TAG}_probe.log 2>&1
python profiles/tools/shapes.py > $O/${TAG}_shapes.log 2>&1
ls -la $O
