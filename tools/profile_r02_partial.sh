O=gpurun_out/r02prof
mkdir -p $O
cap() { local name=$1 rx=$2 skip=$3 cnt=$4; shift 4; "$@" > $O/${name}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o $O/$name "$@" > $O/${name}_ncu.log 2>&1; echo "$name rc=$?"; }
cap gemm_pair_6016x1280x10240 gemm_kernel 2 1 python tools/gemm_one.py 6016 1280 10240 1 1 0 256 3
cap gemm_auto_6016x1280x10240 gemm_kernel 2 1 python tools/gemm_one.py 6016 1280 10240 1 1 0 0 3
cap attn_d40_self attn_kernel 3 3 python tools/attn_one.py full_d40_self 2
