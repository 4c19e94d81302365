python profiles/blend_breakdown.py 2>&1 | tail -3
timeout 300 python -m pytest tests -x -q -m gpu -k "blend or evaluation" 2>&1 | tail -2
python - <<'PY'
import bench, torch, time, sys
sys.argv=['bench.py']
PY
