# End-of-round evidence on one B200: bench lines (both precisions, reference arm), launch lists, ncu captures of the decode-step kernels.
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_1gpu.json 2> gpurun_out/r2h_bench_1gpu.err
python bench.py --steps 5 --warmup 3 --precision bf16 --no-configs > gpurun_out/r2h_bench_1gpu_bf16.json 2> gpurun_out/r2h_bench_1gpu_bf16.err
python tools/run_beam_once.py 5 9472 6 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2h_launches_beam5_9472chunks.csv python tools/run_beam_once.py 5 9472 6 > /dev/null 2>&1
python tools/run_beam_once.py 1 9472 6 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2h_launches_beam1_9472chunks.csv python tools/run_beam_once.py 1 9472 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persistent_kernel -s 4 -c 1 -o gpurun_out/r2h_cell_gemm python tools/run_beam_once.py 5 9472 6 > /dev/null 2>&1
echo done
