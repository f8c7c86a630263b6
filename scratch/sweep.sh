for mb in 4 8 16; do for pc in 6 12; do
echo "max_batch=$mb plane_chunk=$pc"; FVFI_MAX_BATCH=$mb FVFI_PLANE_CHUNK=$pc python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
done; done
