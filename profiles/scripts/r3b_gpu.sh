set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r3b_pytest.log 2>&1; tail -2 $O/r3b_pytest.log
timeout 420 python bench.py --stages > $O/r3b_bench.json 2> $O/r3b_bench.err; tail -c 200 $O/r3b_bench.err
timeout 120 python bench.py --steps 2 --warmup 1 --lean > $O/r3b_lean.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r3b_launches.csv python bench.py --steps 2 --warmup 1 --lean > $O/r3b_ncu_l.log 2>&1
timeout 60 python tools/label_probe.py > $O/r3b_label_probe.log 2>&1 && \
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"pm_label_stream|pm_label_finalize" -s 4 -c 2 -o $O/r3b_label -f python tools/label_probe.py > $O/r3b_ncu_lab.log 2>&1
python profiles/summarise.py kernel $O/r3b_label.ncu-rep > $O/r3b_label.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r3b_smoke.log 2>&1; tail -1 $O/r3b_smoke.log
