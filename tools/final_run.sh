set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2j_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_r2j_c3_1gpu.json 2> gpurun_out/bench_r2j.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2j_reference_arm.json 2> gpurun_out/bench_r2j_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2j_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --subs none > gpurun_out/r2j_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -c 1 -o gpurun_out/gram_r2j -f python bench.py --steps 1 --warmup 0 --samples 1000000 --no-cpu-baseline --no-e2e --subs none > gpurun_out/r2j_ncu_full.log 2>&1
tail -3 gpurun_out/r2j_gputests.txt; cat gpurun_out/r2j_smoke.txt | tail -1; tail -c 600 gpurun_out/bench_r2j_reference_arm.json
