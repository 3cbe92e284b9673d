#!/bin/bash
# One GPU call that checks the whole tree (used during development: bash tools/gpu_full_check.sh TAG under gpurun).
TAG=${1:-chk}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_full.log 2>&1
head -c 6000 gpurun_out/${TAG}_pytest_full.log > gpurun_out/${TAG}_pytest_head.log
tail -60 gpurun_out/${TAG}_pytest_full.log > gpurun_out/${TAG}_pytest.log
grep -n "Error" gpurun_out/${TAG}_pytest_full.log | sort | uniq -c | sort -rn | head -20 > gpurun_out/${TAG}_pytest_errors.log
rm -f gpurun_out/${TAG}_pytest_full.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
timeout 300 python bench.py --config 1 > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err
timeout 300 python bench.py --config 3 > gpurun_out/${TAG}_bench3.json 2> gpurun_out/${TAG}_bench3.err
timeout 400 python bench.py --config 4 > gpurun_out/${TAG}_bench4.json 2> gpurun_out/${TAG}_bench4.err
tail -6 gpurun_out/${TAG}_pytest.log
