#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1
timeout 300 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_stft.py tests/test_gpu_bsseval.py -m gpu -q 2>&1 | tail -4 > $O/gpu_tests_final.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/bench_final_N1.json 2> $O/bench_final_N1.err
echo final done
