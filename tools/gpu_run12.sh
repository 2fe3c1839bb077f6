#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_sconv.py tests/test_gpu_ild.py -m gpu -q -x --durations=3 2>&1 | tail -8 > $O/gpu_tests12.log
timeout 500 python bench.py --config bach10 --steps 5 --no-cpu-baseline --traffic off > $O/bench12_bach10.json 2> $O/bench12_bach10.err
timeout 500 python bench.py --config bach10_score --steps 5 --no-cpu-baseline --traffic off > $O/bench12_score.json 2> $O/bench12_score.err
DCS_DEBUG_TMA_PERSIST=1 timeout 500 python bench.py --config ikala --steps 5 --no-cpu-baseline --traffic off > $O/bench12_ikala.json 2> $O/bench12_ikala.err
echo run12 done
