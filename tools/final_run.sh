# What produced the final round-2 evidence under profiles/ (run through gpurun on one B200, in pieces -- the box time
# of a round is limited):
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/gputests_r2n.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_r2m_c3_1gpu.json 2> gpurun_out/bench_r2m.err
timeout 100 python bench.py --workload c2 --steps 10 --warmup 3 --subs none --no-cpu-baseline > gpurun_out/bench_r2n_c2_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2j_reference_arm.json 2> gpurun_out/bench_r2j_ref.err
# launch list and full capture of the C3 step (kernels unchanged since r2j)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2j_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --subs none > gpurun_out/r2j_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -c 1 -o gpurun_out/gram_r2j -f python bench.py --steps 1 --warmup 0 --samples 1000000 --no-cpu-baseline --no-e2e --subs none > gpurun_out/r2j_ncu_full.log 2>&1
# single-pair kernel of C2
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gram_pow_kernel -c 1 -o gpurun_out/gram_c2_r2m -f python tools/gram_c2_time.py 2 > gpurun_out/r2m_ncu_c2.log 2>&1
python tools/ncu_summary.py gpurun_out/gram_c2_r2m.ncu-rep > profiles/gram_c2_r2m_ncu_summary.txt
# 2-GPU NCCL tests (gpurun --gpus 2)
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "two_gpus or torchrun" > gpurun_out/gputests_r2k_2gpu.txt 2>&1
