#!/bin/bash
# Round-end evidence on ONE B200 (run through gpurun): GPU tests, bench line, reference arm, ncu launch list and one
# `--set full` capture of the three tcgen05 kernels.  Outputs land in gpurun_out/final_*.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/final_pytest.txt 2>&1
tail -2 gpurun_out/final_pytest.txt
python bench.py --steps 100 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/final_ncu1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rnn_.*_tc|wgrad_tc" -s 9 -c 3 -f -o gpurun_out/final_full \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/final_ncu2.err
ls -la gpurun_out/ | head -30
