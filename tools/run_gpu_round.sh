set -x
for i in 1 2; do
SRGAN_DBG_NO_SKIP_FUSE=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-140 | sed 's/^/NOFUSE /'
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-140 | sed 's/^/FUSE /'
done
