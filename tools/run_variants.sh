cd /root/repo
L=deeppreconditioning_b200/lib
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pack or pcg or spmv" 2>&1 | tail -2
run() { echo "== $1 $2"; env $2 DPCG_LIB=$PWD/$L/libdpcg_$1.so python tools/gpu_pack_ab.py --systems 64 --reps 1 $3 2>&1 | grep -v Warning | grep "solves/s"; }
echo "== default (no windows 7680x2, APPLY1 reordered, top-up at release)"; python tools/gpu_pack_ab.py --systems 64 --reps 1 2>&1 | grep "solves/s\|bitwise"
run notopup "" 
echo "=== trace"; DPCG_LIB=$PWD/$L/libdpcg_tr_v1.so python tools/trace_pipe.py 2>&1 | grep -v Warn
