#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
rm -f $O/parity_r2.jsonl
timeout 900 python -m pytest tests/test_gpu_sconv.py tests/test_gpu_gemm.py tests/test_gpu_dsd.py tests/test_gpu_ild.py "tests/test_gpu_fullsize.py::test_bach10_10s_frame4096" "tests/test_gpu_fullsize.py::test_score_informed_10s_frame4096" -m gpu -q --durations=4 2>&1 | tail -12 > $O/gpu_tests20.log
timeout 500 python bench.py --config bach10 --steps 5 > $O/bench20_bach10.json 2> $O/bench20_bach10.err
timeout 500 python bench.py --config bach10_score --steps 5 > $O/bench20_score.json 2> $O/bench20_score.err
DCS_DEBUG_TMA_MASK=31 timeout 500 python bench.py --config bach10_score --steps 5 --no-cpu-baseline --traffic off > $O/bench20_score_nowin.json 2> $O/bench20_score_nowin.err
echo run20 done
