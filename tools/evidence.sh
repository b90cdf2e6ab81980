# Round evidence recipe (run under gpurun on one B200): smoke, bench, ncu launch list, ncu --set full of the
# headline kernel, of one config-5 pass and of config 3 at B=128.  Every ncu command runs after the same
# command has exited 0 without ncu.  Outputs land in gpurun_out/ (copy what is to be judged into profiles/).
set -x
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r1f}
python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1
python bench.py > gpurun_out/${T}_bench.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_search -s 3 -c 2 -o gpurun_out/${T}_fused -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu2.log 2>&1
ncu -i gpurun_out/${T}_fused.ncu-rep --page raw --csv > gpurun_out/${T}_fused_raw.csv 2>/dev/null
python tools/one_case.py 12500000 fp16 256 100 3 > gpurun_out/${T}_c5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_search -s 2 -c 1 -o gpurun_out/${T}_c5 -f python tools/one_case.py 12500000 fp16 256 100 3 > gpurun_out/${T}_ncu3.log 2>&1
ncu -i gpurun_out/${T}_c5.ncu-rep --page raw --csv > gpurun_out/${T}_c5_raw.csv 2>/dev/null
python tools/one_case.py 1000000 fp32 128 10 3 > gpurun_out/${T}_c3_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:fused_search -s 1 -c 1 -o gpurun_out/${T}_c3 -f python tools/one_case.py 1000000 fp32 128 10 3 > gpurun_out/${T}_ncu4.log 2>&1
ncu -i gpurun_out/${T}_c3.ncu-rep --page raw --csv > gpurun_out/${T}_c3_raw.csv 2>/dev/null
python tools/one_case.py 1000000 fp32 32 10 3 > gpurun_out/${T}_c3b32_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:fused_search -s 1 -c 1 -o gpurun_out/${T}_c3b32 -f python tools/one_case.py 1000000 fp32 32 10 3 > gpurun_out/${T}_ncu5.log 2>&1
ncu -i gpurun_out/${T}_c3b32.ncu-rep --page raw --csv > gpurun_out/${T}_c3b32_raw.csv 2>/dev/null
for c in 1 2 3 5; do python bench.py --config $c > gpurun_out/${T}_bench_c$c.log 2>&1; done
rm -f gpurun_out/${T}_c3b32.ncu-rep gpurun_out/${T}_c3.ncu-rep gpurun_out/${T}_fused.ncu-rep   # (the config-5 report stays: source page)
tail -3 gpurun_out/${T}_smoke.log; tail -1 gpurun_out/${T}_bench.log; cat gpurun_out/${T}_c5_plain.log
