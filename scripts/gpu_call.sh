# one GPU-box call of round 2 (tag r02w): full GPU test suite on the shipped library, decode on the 1 Mb contig shape,
# k_fb2 timing of the transposed s->M table variants, parity tests on the variant
T=${TAG:-r02w}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_gpu.log
python scripts/decode_compare.py 300 5000 1000000 100 > gpurun_out/${T}_decode_compare_1mb.json 2> gpurun_out/${T}_decode_compare_1mb.err
for lib in nanopore_b200/libphmm_sm100.so build/libphmm_tmt.so build/libphmm_tmt5.so nanopore_b200/libphmm_sm100.so build/libphmm_tmt.so; do
  echo "== $lib" >> gpurun_out/${T}_tune.log
  TUNE_LIB=$lib REPS=3 python scripts/tune.py 2368 "" >> gpurun_out/${T}_tune.log 2>&1
done
PHMM_LIB=$PWD/build/libphmm_tmt.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${T}_pytest_tmt.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_tmt.log
true
