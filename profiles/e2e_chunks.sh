for ch in 16 24 32 48; do
  GLSB_HOST_CHUNKS=$ch python bench.py --steps 5 --warmup 3 --no-cpu-baseline --time-step-refinements -1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks', $ch, 'e2e', round(d['e2e']['value'],3), 'value', round(d['value'],2))"
done
