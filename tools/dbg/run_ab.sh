# scratch A/B driver for gpurun (tuning only; numbers quoted in DESIGN.md come from bench.py / tools/kbench.py runs)
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
