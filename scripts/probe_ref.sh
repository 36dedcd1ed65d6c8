#!/bin/bash
# First action of every gpurun: is the reference stack reachable on this box?
# If it is, dump the golden vectors from the real jax + gymnax into gpurun_out/golden_ref/ (to be committed under
# tests/golden/).  Never fails the calling command.
mkdir -p gpurun_out
{
  echo "== probe $(date -u +%FT%TZ)"
  ls -la baseline/_ref 2>&1 | head -5
  PYTHONPATH=baseline/_ref:$PYTHONPATH python - <<'PY'
import importlib
for m in ("jax", "jaxlib", "flax", "optax", "gymnax", "chex"):
    try:
        mod = importlib.import_module(m)
        print(m, "OK", getattr(mod, "__version__", "?"))
    except Exception as e:
        print(m, "MISSING", type(e).__name__)
PY
  if PYTHONPATH=baseline/_ref:$PYTHONPATH python -c "import jax, gymnax" 2>/dev/null; then
    PYTHONPATH=baseline/_ref:$PYTHONPATH timeout 900 python tests/golden/make_golden_from_ref.py --out gpurun_out/golden_ref
  fi
} > gpurun_out/probe_ref.log 2>&1
tail -8 gpurun_out/probe_ref.log
exit 0
