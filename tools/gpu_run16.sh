#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_dsd.py tests/test_gpu_sconv.py tests/test_gpu_ild.py -m gpu -q -x --durations=3 2>&1 | tail -8 > $O/gpu_tests16.log
timeout 400 python bench.py --steps 10 --clips 8 --no-cpu-baseline --traffic off > $O/bench16_N1.json 2> $O/bench16_N1.err
timeout 500 python bench.py --config bach10 --steps 5 --no-cpu-baseline --traffic off > $O/bench16_bach10.json 2> $O/bench16_bach10.err
timeout 500 python bench.py --config bach10_score --steps 5 --no-cpu-baseline --traffic off > $O/bench16_score.json 2> $O/bench16_score.err
timeout 500 python bench.py --config ikala --steps 5 --no-cpu-baseline --traffic off > $O/bench16_ikala.json 2> $O/bench16_ikala.err
echo run16 done
