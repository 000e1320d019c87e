for TD in 4 5 6; do SFFTB_BENCH_TILE_DEPTH=$TD python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import sys, json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); c=d['config4']; print('TD', c['tiles_in_flight'], c['value'], c['ms_per_tile_per_gpu'])"; done
