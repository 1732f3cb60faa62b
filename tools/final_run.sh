#!/bin/bash
# The round's closing run on one GPU: the whole GPU suite, the bench line of every single-GPU workload, the reference
# arm, then the ncu launch list of the default bench command and one full capture of one isosurface's kernels.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
for w in cfg1 cfg3 cfg5; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/fin_$w.json 2> gpurun_out/fin_$w.err; tail -c 200 gpurun_out/fin_$w.err; done
python bench.py --steps 20 --warmup 3 > gpurun_out/fin_cfg2.json 2> gpurun_out/fin_cfg2.err; tail -c 200 gpurun_out/fin_cfg2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_ref.json 2> gpurun_out/fin_ref.err
for w in cfg1 cfg2 cfg3 cfg5 ref; do head -c 300 gpurun_out/fin_$w.json; echo; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_count2|k_emit_cells|k_emit_vertices|k_classify" -s 1 -c 6 -f -o gpurun_out/r2f_prof python tools/profile_run.py 512 1 > gpurun_out/ncu_f2.log 2>&1
tail -2 gpurun_out/ncu_f2.log
