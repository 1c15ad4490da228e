mkdir -p gpurun_out
python tools/matvec_bench.py > gpurun_out/r02_matvec_row.log 2>&1; tail -1 gpurun_out/r02_matvec_row.log | cut -c1-1200
python tools/matvec_bench.py --layout transposed > gpurun_out/r02_matvec_trans.log 2>&1; tail -1 gpurun_out/r02_matvec_trans.log | cut -c1-1200
