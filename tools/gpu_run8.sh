#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_dsd.py tests/test_gpu_sconv.py tests/test_gpu_ild.py -m gpu -q -x --durations=3 2>&1 | tail -12 > $O/gpu_tests8.log
for P in 0 4; do
DCS_DEBUG_TMA_PERSIST=$P timeout 400 python bench.py --steps 10 --no-cpu-baseline --traffic off > $O/bench8_N1_p$P.json 2> $O/bench8_N1_p$P.err
DCS_DEBUG_TMA_PERSIST=$P timeout 500 python bench.py --config bach10 --steps 5 --no-cpu-baseline --traffic off > $O/bench8_bach10_p$P.json 2> $O/bench8_bach10_p$P.err
DCS_DEBUG_TMA_PERSIST=$P timeout 500 python bench.py --config ikala --steps 5 --no-cpu-baseline --traffic off > $O/bench8_ikala_p$P.json 2> $O/bench8_ikala_p$P.err
done
echo run8 done
