cd /root/repo
L=deeppreconditioning_b200/lib
run() { echo "== $1 $2"; env $2 DPCG_LIB=$PWD/$L/libdpcg_$1.so timeout 300 python tools/gpu_pack_ab.py --systems 64 --reps 1 $3 2>&1 | grep -v Warning | grep "solves/s\|Error\|error"; }
echo "== default"; timeout 300 python tools/gpu_pack_ab.py --systems 64 --reps 1 --packed-only 2>&1 | grep "solves/s"
for v in ua5 ua5_1_5 ua5_2_5 u2_8 ua2; do run $v "" --packed-only; done
