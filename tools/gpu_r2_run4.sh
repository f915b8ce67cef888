#!/bin/bash
# Round 2, run 4: encoder v2 tuned (acceptance bias, fewer barriers, place loop, LZ table), CRC table variants, foreign decoder timings.
mkdir -p gpurun_out
{
echo "=== pytest parity"; S=$(date +%s); timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -12; echo "elapsed $(( $(date +%s) - S )) s"
for crc in 0; do
echo "=== phase profile v2 crc=$crc"; MFCAS_ENC_CRC=$crc MFCAS_PHASE_TIMING=1 timeout 600 python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep "phase\] deflate" | tail -2 | head -1
echo "=== bench C2 crc=$crc"; MFCAS_ENC_CRC=$crc timeout 1200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c2v2.err | tee gpurun_out/bench_c2_crc$crc.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['config']['packed_ratio'], d['kernel_ms_per_step'], d['e2e'])"; tail -3 gpurun_out/c2v2.err
done
echo "=== ncu deflate2"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:mf_deflate2_kernel -s 1 -c 1 -o gpurun_out/r02_deflate2b -f python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 > gpurun_out/ncu_d2.log 2>&1; tail -1 gpurun_out/ncu_d2.log | cut -c1-200
echo "=== bench C2 3072 blobs/step"; timeout 1200 python bench.py --blobs-per-step 3072 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c2b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"; tail -3 gpurun_out/c2b.err
echo "=== foreign decode launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mf_ -c 200 --csv --log-file gpurun_out/r02_foreign_launches.csv python tools/foreign_decode.py > gpurun_out/ncu_foreign.log 2>&1; tail -3 gpurun_out/ncu_foreign.log
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_foreign_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); iu=h.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[hdr+1:]:
    try: v=float(r[iv].replace(',',''))
    except: continue
    if r[iu]=='ns': v/=1e6
    elif r[iu]=='us': v/=1e3
    elif r[iu]=='s': v*=1e3
    k=r[ik].split('(')[0]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
for k,(n,t) in agg.items(): print("%-28s launches %3d total %.3f ms"%(k,n,t))
PY
echo "=== C3 timeline"; timeout 600 python tools/c3_small_blobs.py 2>&1 | tail -4
} > gpurun_out/r2_run4.log 2>&1
tail -c 6000 gpurun_out/r2_run4.log
