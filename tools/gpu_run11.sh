#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_dsd.py -m gpu -q -x --durations=3 2>&1 | tail -15 > $O/gpu_tests11.log
timeout 300 python -m pytest tests/test_gpu_sconv.py -m gpu -q -x -k "small or full_size or score" 2>&1 | tail -8 >> $O/gpu_tests11.log
for A in 1 0; do
DCS_DEBUG_TMA_ATM=$A timeout 400 python bench.py --steps 10 --clips 8 --no-cpu-baseline --traffic off > $O/bench11_N1_atm$A.json 2> $O/bench11_N1_atm$A.err
DCS_DEBUG_TMA_ATM=$A timeout 500 python bench.py --config bach10 --steps 5 --no-cpu-baseline --traffic off > $O/bench11_bach10_atm$A.json 2> $O/bench11_bach10_atm$A.err
done
timeout 400 python bench.py --steps 10 --no-cpu-baseline --traffic off > $O/bench11_N1_clips32.json 2> $O/bench11_N1_clips32.err
echo run11 done
