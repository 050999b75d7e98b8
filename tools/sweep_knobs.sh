for l in 16 32; do for g in 8 16; do for pm in 3 4; do for sb in 8 24; do
r=$(CAAA_LANES=$l CAAA_GROUP=$g CAAA_POOL_MULT=$pm CAAA_SWITCH_BELOW=$sb timeout 100 python bench.py --playouts 1048576 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; print(int(json.loads(sys.stdin.read().strip().splitlines()[-1])['value']))")
echo "L=$l G=$g PM=$pm SB=$sb $r"
done; done; done; done
