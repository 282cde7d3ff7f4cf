set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 400 gpurun_out/r02_bench.json; tail -2 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/r02_bench_reference.json 2>/dev/null; tail -c 600 gpurun_out/r02_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fsk_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:fsk_demod_pipe -s 2 -c 1 -o gpurun_out/r02_pipe python profiles/sweep.py 20 pipe:256 2>&1 | tail -1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:fsk_wpc -s 2 -c 1 -o gpurun_out/r02_wpc_c python profiles/sweep.py 16 wpc:2220 2>&1 | tail -1
