mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/t15_tests.log 2>&1; echo "rc=$?" >> gpurun_out/t15_tests.log); tail -3 gpurun_out/t15_tests.log
timeout 100 python tools/bench_tc.py gemm 2>&1 | cut -c1-210
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t15_bench_c2.json 2>gpurun_out/t15_bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/t15_bench_c2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['host_enqueue_ms_per_step'], d['roofline']['stage_ms_all'])"
SBR_SCAN_NO_COOP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_scan_.*_kernel" -s 4 -c 4 -f -o gpurun_out/r2_full_c3_scans python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --min-timed-s 0 > /dev/null 2> gpurun_out/r2_ncu5.err; tail -3 gpurun_out/r2_ncu5.err
SBR_SCAN_NO_COOP=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c3.csv python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --min-timed-s 0 > /dev/null 2> gpurun_out/r2_ncu6.err; grep -c tc_scan gpurun_out/r2_launches_c3.csv
