set -x
cd $GRAFT_REPO_ROOT
python __graft_entry__.py --smoke > gpurun_out/r1f_smoke.log 2>&1
python bench.py > gpurun_out/r1f_bench.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_search -s 3 -c 2 -o gpurun_out/r1f_fused -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_ncu2.log 2>&1
ncu -i gpurun_out/r1f_fused.ncu-rep --page raw --csv > gpurun_out/r1f_fused_raw.csv 2>/dev/null
python tools/one_case.py 12500000 fp16 256 100 3 > gpurun_out/r1f_c5_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:fused_search -s 2 -c 2 -o gpurun_out/r1f_c5 -f python tools/one_case.py 12500000 fp16 256 100 3 > gpurun_out/r1f_ncu3.log 2>&1
ncu -i gpurun_out/r1f_c5.ncu-rep --page raw --csv > gpurun_out/r1f_c5_raw.csv 2>/dev/null
python tools/one_case.py 1000000 fp32 128 10 3 > gpurun_out/r1f_c3_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:fused_search -s 1 -c 1 -o gpurun_out/r1f_c3 -f python tools/one_case.py 1000000 fp32 128 10 3 > gpurun_out/r1f_ncu4.log 2>&1
ncu -i gpurun_out/r1f_c3.ncu-rep --page raw --csv > gpurun_out/r1f_c3_raw.csv 2>/dev/null
rm -f gpurun_out/r1f_c5.ncu-rep gpurun_out/r1f_c3.ncu-rep
ls -la gpurun_out | tail -20
cat gpurun_out/r1f_smoke.log | tail -3; cat gpurun_out/r1f_bench.log | tail -1
