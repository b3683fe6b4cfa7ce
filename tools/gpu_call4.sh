set -x
mkdir -p gpurun_out
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_v2_n1.json 2> gpurun_out/r02_bench_v2_n1.err; echo bench rc=$?
timeout 400 python tools/sensitivity.py > gpurun_out/r02_sensitivity_n1.jsonl 2> gpurun_out/r02_sensitivity_n1.err; echo sens rc=$?; cat gpurun_out/r02_sensitivity_n1.jsonl; tail -c 400 gpurun_out/r02_sensitivity_n1.err
timeout 300 python tools/run_config4_sharded.py --grna 20000 > gpurun_out/r02_config4_n1.jsonl 2> gpurun_out/r02_config4_n1.err; echo c4 rc=$?; cat gpurun_out/r02_config4_n1.jsonl; tail -c 400 gpurun_out/r02_config4_n1.err
B="python bench.py --steps 2 --warmup 1 --no-other-modes --no-e2e --no-checks --no-cpu-baseline"
$B > gpurun_out/ncu_plain1.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_bench_launches_raw.csv $B > gpurun_out/ncu1.log 2>&1; echo launches rc=$?
S="python bench.py --genome-mb 17 --steps 1 --warmup 1 --no-other-modes --no-e2e --no-checks --no-cpu-baseline"
$S > gpurun_out/ncu_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:screen_pairs -s 1 -c 1 -o gpurun_out/r02_pairs_prof $S > gpurun_out/ncu2.log 2>&1; echo pairs rc=$?
G="python bench.py --gapped --grna 1024 --genome-mb 17 --steps 1 --warmup 1 --no-e2e --no-checks"
$G > gpurun_out/ncu_plain3.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gapped_rows -s 2 -c 1 -o gpurun_out/r02_rows_prof $G > gpurun_out/ncu3.log 2>&1; echo rows rc=$?
cd tools/k6
./k6_proto 256 > ../../gpurun_out/ncu_plain4.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k6_kernelILi8ELi0E -s 2 -c 1 -o ../../gpurun_out/r02_k6_prof ./k6_proto 256 > ../../gpurun_out/ncu4.log 2>&1; echo k6 rc=$?
cd ../..
ls -la gpurun_out/*.ncu-rep
