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
