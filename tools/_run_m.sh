mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_m_pytest.log
tail -3 gpurun_out/r2_m_pytest.log
timeout 900 compute-sanitizer --tool memcheck --kernel-regex kne=decoder_tc_kernel python tools/dec_once.py 20 fp16x3 1 > gpurun_out/r2_m_memcheck.log 2>&1; tail -4 gpurun_out/r2_m_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --kernel-regex kne=decoder_tc_kernel python tools/dec_once.py 12 fp16x3 1 > gpurun_out/r2_m_racecheck.log 2>&1; tail -4 gpurun_out/r2_m_racecheck.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
