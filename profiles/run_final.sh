set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 300 gpurun_out/r02_bench.json; tail -2 gpurun_out/r02_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fsk_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-sweep > /dev/null 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:fsk_duo -s 2 -c 1 -o gpurun_out/r02_duo_kernel python profiles/sweep.py 20 duo:256 2>&1 | tail -1
