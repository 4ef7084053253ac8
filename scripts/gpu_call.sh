# final validation of a round: smoke, bench, launch list, ncu captures, EM at the metric's band (GPU box)
T=${TAG:-r02v}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --reads 2960 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_w -s 1 -c 1 -o gpurun_out/${T}_prof_decode_w -f python bench.py --reads 2368 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_prof_decode_w.log 2>&1
EM_BAND=50 EM_CPU_SAMPLE=16 python tests/tools/em_bench.py 2000 3 > gpurun_out/${T}_em_band50.json 2> gpurun_out/${T}_em_band50.err
true
