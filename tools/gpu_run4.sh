#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
rm -f $O/parity_r2.jsonl
timeout 1500 python -m pytest tests/test_gpu_sconv.py tests/test_gpu_ild.py tests/test_gpu_dropin.py tests/test_gpu_bsseval.py "tests/test_gpu_fullsize.py::test_bach10_10s_frame4096" "tests/test_gpu_fullsize.py::test_ikala_pooled_10s_with_silence" "tests/test_gpu_fullsize.py::test_score_informed_10s_frame4096" -m gpu -q --durations=8 2>&1 | tail -40 > $O/gpu_tests4.log
for c in bach10 bach10_score ikala; do
  timeout 500 python bench.py --config $c --steps 5 --no-cpu-baseline > $O/bench4_$c.json 2> $O/bench4_$c.err
done
DCS_DEBUG_SIMT_GEMM=1 timeout 600 python -m pytest tests/test_gpu_sconv.py -m gpu -q -k "small or nopool or score" 2>&1 | tail -5 > $O/gpu_tests4_simt.log
echo run4 done
