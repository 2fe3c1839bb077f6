#!/bin/bash
# one GPU-box call: host facts, the whole GPU test suite (incl. the full-size strict parity cases), bench lines for every
# configuration, optionally the reference arm as the driver calls it, a launch list and a full ncu capture of the top kernels
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
rm -f $O/parity_r2.jsonl
{ echo "nproc $(nproc)"; echo "cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; free -g | head -2; nvidia-smi topo -m; } > $O/host.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -40 > $O/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 500 python bench.py --steps 20 --warmup 5 > $O/bench_N1.json 2> $O/bench_N1.err
for c in dsd1024 bach10 bach10_score ikala; do
  timeout 500 python bench.py --config $c --steps 5 > $O/bench_$c.json 2> $O/bench_$c.err
done
if [ "$1" = "ref" ]; then
  ( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches.csv python bench.py --traffic-probe --clips 1 --device-streams 1 --no-numa > /dev/null 2> $O/ncu_launch.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"istft_reg|dsd_mask_tc|stft_reg_kernel|gemm_tma" -s 12 -c 12 -o $O/r2_prof python bench.py --traffic-probe --clips 1 --device-streams 1 --no-numa > /dev/null 2> $O/ncu_full.err
echo suite done
