set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02_tests_multi.log; tail -5 gpurun_out/r02_tests_multi.log
timeout 300 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_v2_n8.json 2> gpurun_out/r02_bench_v2_n8.err; echo n8 rc=$?; tail -c 300 gpurun_out/r02_bench_v2_n8.err
timeout 300 $TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_v2_n2.json 2> gpurun_out/r02_bench_v2_n2.err; echo n2 rc=$?; tail -c 300 gpurun_out/r02_bench_v2_n2.err
timeout 300 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --gapped --grna 10000 --steps 2 --warmup 1 > gpurun_out/r02_bench_gapped_v2_n8.json 2> gpurun_out/r02_bench_gapped_v2_n8.err; echo gapped8 rc=$?; tail -c 300 gpurun_out/r02_bench_gapped_v2_n8.err
timeout 400 $TR --nproc-per-node 8 --master-port 29614 tools/run_config4_sharded.py --grna 20000 --grna 1000000 > gpurun_out/r02_config4_n8.jsonl 2> gpurun_out/r02_config4_n8.err; echo c4 rc=$?; cat gpurun_out/r02_config4_n8.jsonl; tail -c 400 gpurun_out/r02_config4_n8.err
timeout 120 python bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_arm.json 2>/dev/null; cat gpurun_out/r02_bench_ref_arm.json | head -c 600
