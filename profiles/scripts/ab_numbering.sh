for nb in morton engine blockMesh; do
python bench.py --cells 200 --gas air5 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --numbering $nb 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('$nb', round(d['ms_per_step'],3), {a:round(b,3) for a,b in k.items()})"
done
