#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_dsd.py tests/test_gpu_sconv.py tests/test_gpu_ild.py -m gpu -q -x --durations=5 2>&1 | tail -25 > $O/gpu_tests6.log
timeout 400 python bench.py --steps 10 --no-cpu-baseline --traffic off > $O/bench6_N1.json 2> $O/bench6_N1.err
DCS_DEBUG_TMA_PERSIST=0 timeout 400 python bench.py --steps 10 --no-cpu-baseline --traffic off > $O/bench6_N1_np.json 2> $O/bench6_N1_np.err
timeout 500 python bench.py --config bach10 --steps 5 --no-cpu-baseline --traffic off > $O/bench6_bach10.json 2> $O/bench6_bach10.err
timeout 500 python bench.py --config ikala --steps 5 --no-cpu-baseline --traffic off > $O/bench6_ikala.json 2> $O/bench6_ikala.err
timeout 500 python bench.py --config bach10_score --steps 5 --no-cpu-baseline --traffic off > $O/bench6_score.json 2> $O/bench6_score.err
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_sconv.py -m gpu -q -k "small or nopool" 2>&1 | tail -8 > $O/memcheck6.log
echo run6 done
