O=gpurun_out/r02prof
mkdir -p $O
BENCH="python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-sampling --no-rvq --no-full-step"
$BENCH > $O/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 5500 -c 30000 --csv --log-file $O/launches_all.csv $BENCH > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"; grep -c add_noise $O/launches_all.csv; wc -l $O/launches_all.csv
cuobjdump -sass prompt_tts_b200/libpt_b200.so | grep -oE "UTCHMMA[.A-Z0-9_]*|LDTM[.A-Za-z0-9_]*|STTM[.A-Za-z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|REDG[.A-Za-z0-9_]*|HMMA[.A-Z0-9_]*" | sort | uniq -c > $O/sass_mnemonics.txt
