for mb in 4 8 16 32 64; do PEEB_CHUNK_MB=$mb python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk_mb', $mb, 'e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],2))"; done
