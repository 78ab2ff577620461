set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2v_pytest.log 2>&1; tail -3 $O/r2v_pytest.log
timeout 420 python bench.py --stages > $O/r2v_bench.json 2> $O/r2v_bench.err; tail -c 300 $O/r2v_bench.err
timeout 120 python bench.py --steps 2 --warmup 1 --lean > $O/r2v_lean.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2v_launches.csv python bench.py --steps 2 --warmup 1 --lean > $O/r2v_ncu_l.log 2>&1
timeout 60 python tools/label_probe.py > $O/r2v_label_probe.log 2>&1 && \
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"pm_label_stream|pm_label_finalize" -s 4 -c 2 -o $O/r2v_label -f python tools/label_probe.py > $O/r2v_ncu_lab.log 2>&1
python profiles/summarise.py kernel $O/r2v_label.ncu-rep > $O/r2v_label.txt 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2v_bench_reference.json 2> $O/r2v_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2v_smoke.log 2>&1; tail -1 $O/r2v_smoke.log
