set -x
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
for w in cfg1 cfg3 cfg5; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/fin_$w.json 2> gpurun_out/fin_$w.err; tail -c 200 gpurun_out/fin_$w.err; done
python bench.py --steps 20 --warmup 3 > gpurun_out/fin_cfg2.json 2> gpurun_out/fin_cfg2.err; tail -c 200 gpurun_out/fin_cfg2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_ref.json 2> gpurun_out/fin_ref.err
for w in cfg1 cfg2 cfg3 cfg5 ref; do head -c 400 gpurun_out/fin_$w.json; echo; done
