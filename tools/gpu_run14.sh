#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_atm -s 3 -c 1 -o $O/r2_prof_bach10_convT2 python bench.py --config bach10 --traffic-probe --clips 1 --device-streams 1 --no-numa > /dev/null 2> $O/ncu_b10.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_atm -s 1 -c 1 -o $O/r2_prof_dsd_convT2 python bench.py --traffic-probe --clips 1 --device-streams 1 --no-numa > /dev/null 2> $O/ncu_dsd.err
echo run14 done
