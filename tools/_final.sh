set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cat gpurun_out/bench_final.json | cut -c1-1500
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-600
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --pairs 32 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:r3d:: -s 1400 -c 700 --csv --log-file gpurun_out/launches_final.csv python bench.py --pairs 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1; tail -2 gpurun_out/ncu_final.log | cut -c1-300
