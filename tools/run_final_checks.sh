#!/bin/bash
# Round-end checks on one B200 (gpurun): GPU tests, smoke, the bench line + variants, the reference arm, ncu captures.
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_k20.json 2> gpurun_out/r02_final_bench.err
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02_final_bench_k50.json 2>> gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err
for b in 512 4096; do timeout 600 python bench.py --steps 20 --warmup 5 --batch $b --no-cpu-baseline > gpurun_out/r02_final_bench_b$b.json 2>> gpurun_out/r02_final_bench.err; done
timeout 600 python bench.py --steps 20 --warmup 5 --graph rmat --no-cpu-baseline > gpurun_out/r02_final_bench_rmat.json 2>> gpurun_out/r02_final_bench.err
python tools/show_bench.py gpurun_out/r02_final_bench_*.json gpurun_out/r02_final_ref.json
tail -2 gpurun_out/r02_final_pytest.log; tail -1 gpurun_out/r02_final_smoke.log
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$B > gpurun_out/ncu_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_final_launches.csv $B > gpurun_out/ncu1.log 2>&1; echo rc1=$?
$B > gpurun_out/ncu_plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gather_mean_kernel|tc_gemm_kernel|head_rows_kernel" -s 4 -c 8 -o gpurun_out/r02_final_prof -f $B > gpurun_out/ncu2.log 2>&1; echo rc2=$?
# multi-GPU (gpurun --gpus 8): both step counts -- the 48-step run is the one that hung before the wait-rule fix of R2.8
# R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"; A="--gpus 8 --warmup 5 --no-cpu-baseline --no-kernel-profile"
# for k in 20 48; do timeout 300 $R --master-port 2961$((k % 10)) bench.py $A --steps $k > gpurun_out/final_n8_k$k.json 2> gpurun_out/final_n8_k$k.err; done
# timeout 300 $R --master-port 29619 tests/multigpu_dp_check.py 2>&1 | tail -1
