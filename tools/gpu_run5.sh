#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_stft.py tests/test_gpu_dsd.py tests/test_gpu_gemm.py tests/test_gpu_bsseval.py tests/test_gpu_dropin.py -m gpu -q --durations=5 2>&1 | tail -25 > $O/gpu_tests5.log
timeout 400 python bench.py --steps 10 --no-cpu-baseline --traffic off > $O/bench5_N1.json 2> $O/bench5_N1.err
timeout 400 python bench.py --steps 10 --no-cpu-baseline --traffic off --device-streams 4 > $O/bench5_N1_ds4.json 2> $O/bench5_N1_ds4.err
DCS_DEBUG_TMA_PREFETCH=0 timeout 400 python bench.py --steps 10 --no-cpu-baseline --traffic off > $O/bench5_N1_pf0.json 2> $O/bench5_N1_pf0.err
timeout 500 python bench.py --config bach10 --steps 5 --no-cpu-baseline --traffic off > $O/bench5_bach10.json 2> $O/bench5_bach10.err
DCS_DEBUG_TMA_PREFETCH=12 timeout 500 python bench.py --config bach10 --steps 5 --no-cpu-baseline --traffic off > $O/bench5_bach10_pf12.json 2> $O/bench5_bach10_pf12.err
timeout 500 python bench.py --config ikala --steps 5 --no-cpu-baseline --traffic off > $O/bench5_ikala.json 2> $O/bench5_ikala.err
echo run5 done
