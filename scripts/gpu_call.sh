# final GPU-box call of round 2 (tag r02z): the shipped library (s->M table layout 2, launch bound 5, k_decode_w): tests, smoke, bench,
# launch list, one full ncu capture of k_fb2
T=${TAG:-r02z}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --reads 2960 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_b_ncu.log 2>&1
# (this capture ran into the call's time limit in round 2 and the box was returned wedged: keep a shorter timeout of its own on it)
SHIPPED_LIB=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_fb2 -s 1 -c 1 -o gpurun_out/${T}_prof_fb2 -f python scripts/tune.py 740 "" > gpurun_out/${T}_prof_fb2.log 2>&1
true
