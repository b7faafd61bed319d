set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err
python bench.py --steps 5 --warmup 3 --precision bf16 --no-configs > gpurun_out/r2f_bench_1gpu_bf16.json 2> gpurun_out/r2f_bench_1gpu_bf16.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference_arm.json 2> gpurun_out/r2f_ref.err
python tools/run_beam_once.py 5 9472 6 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_beam5_9472chunks.csv python tools/run_beam_once.py 5 9472 6 > /dev/null 2>&1
python tools/run_beam_once.py 1 9472 6 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_beam1_9472chunks.csv python tools/run_beam_once.py 1 9472 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -o gpurun_out/r2f_att_tc python tools/run_beam_once.py 5 9472 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fc_search_kernel -s 2 -c 1 -o gpurun_out/r2f_fc_search python tools/run_beam_once.py 5 9472 6 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persistent_kernel -s 4 -c 1 -o gpurun_out/r2f_cell_gemm python tools/run_beam_once.py 5 9472 6 > /dev/null 2>&1
echo done
