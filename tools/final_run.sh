set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r01b_bench_default.json 2> gpurun_out/r01b_bench_default.err; echo rc=$?
python bench.py --impl reference > gpurun_out/r01b_bench_reference.json 2> gpurun_out/r01b_bench_reference.err; echo rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r01b_pre_ncu.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r01b_ncu1.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:'pvalues_kernel|hist_pairs|bh_rank|bh_mask|fit_kernel' -s 15 -c 5 -f -o gpurun_out/r01b_full python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r01b_ncu2.log 2>&1; echo rc=$?
ls -la gpurun_out/r01b_*
