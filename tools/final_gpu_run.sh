# Round-end check on a B200 box: GPU tests, smoke, both bench arms; optionally (LAUNCHES=1) the ncu launch list
# of the benched step.  Outputs under gpurun_out/final/ (copied into profiles/ by hand).
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/final/gpu_tests.txt 2>&1; tail -3 gpurun_out/final/gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.txt 2>&1; tail -1 gpurun_out/final/smoke.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; tail -c 400 gpurun_out/final/bench_ref.json
python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; tail -c 300 gpurun_out/final/bench.json
if [ -n "$LAUNCHES" ]; then
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 1944 -c 860 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/final/ncu.log 2>&1
wc -l gpurun_out/final/launches.csv
fi
