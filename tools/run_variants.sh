cd /root/repo
L=deeppreconditioning_b200/lib
run() { echo "== $1 $2"; env $2 DPCG_LIB=$PWD/$L/libdpcg_$1.so python tools/gpu_pack_ab.py --systems 64 --reps 1 $3 2>&1 | grep -v Warning | grep "solves/s\|bitwise\|Error\|error"; }
run rr ""
run rr_top ""
run rr3_top "" --packed-only
echo "=== trace rr_top"; DPCG_LIB=$PWD/$L/libdpcg_tr_rr_top.so python tools/trace_pipe.py 2>&1 | grep -v Warn
