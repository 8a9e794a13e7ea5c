#!/bin/bash
# the driver's round-end sequence on one GPU: tests, smoke, reference arm, our arm (defaults)
TAG=${1:-contract}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/${TAG}_smoke.log
( time python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 ) > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"; cat gpurun_out/${TAG}_ref.json; tail -4 gpurun_out/${TAG}_ref.err
( time python bench.py --gpus 1 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -4 gpurun_out/${TAG}_bench.err
