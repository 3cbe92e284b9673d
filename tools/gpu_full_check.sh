python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/r02h_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02h_smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/r02h_bench2.json 2> gpurun_out/r02h_bench2.err
WB_TC_FUSE_SCATTER=3 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02h_bench2_fuse3.json 2> gpurun_out/r02h_bench2_fuse3.err
WB_TC_FWD_PIPE=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02h_bench2_pipe3.json 2> gpurun_out/r02h_bench2_pipe3.err
WB_TC_FWD_PIPE=1 WB_TC_FWD_CTAS=2 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02h_bench2_pipe2.json 2> gpurun_out/r02h_bench2_pipe2.err
timeout 300 python bench.py --config 3 > gpurun_out/r02h_bench3.json 2> gpurun_out/r02h_bench3.err
timeout 400 python bench.py --config 4 > gpurun_out/r02h_bench4.json 2> gpurun_out/r02h_bench4.err
tail -6 gpurun_out/r02h_pytest.log
