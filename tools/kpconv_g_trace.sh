set -x
cd superpoints_registration_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -DSPR_G_TRACE -c kpconv_g.cu -o build/kpconv_g.o && nvcc -shared -o libspr_b200.so build/*.o; echo "build rc=$?"
cd ../..
python tools/kpconv_g_trace.py 32 2>&1 | tail -20
python tools/kpconv_g_trace.py 128 2>&1 | tail -20
