#!/bin/bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with gpurun) for the
# in-situ arm (BASELINE.json configs[4]): the vendored SpeechBrain via pip from a /tmp copy (the build writes into its
# source tree and /root/reference is read-only), plus the recipe's own files (training script, models/, utils.py,
# hparams/) verbatim under baseline/_ref/recipe.  Authoring container only; nothing here is committed.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
rm -rf "$ROOT/baseline/_ref" /tmp/tsasr_sbcopy
mkdir -p "$ROOT/baseline/_ref" /tmp/tsasr_sbcopy
cp -r "$REF/vendor/speechbrain/." /tmp/tsasr_sbcopy/
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" /tmp/tsasr_sbcopy
rm -rf "$ROOT/baseline/_ref/tests"          # the wheel drops a top-level "tests" package that would shadow ours
mkdir -p "$ROOT/baseline/_ref/recipe"
cp "$REF/train_librispeechmix_scratch.py" "$REF/utils.py" "$ROOT/baseline/_ref/recipe/"
cp -r "$REF/models" "$REF/hparams" "$ROOT/baseline/_ref/recipe/"
rm -rf /tmp/tsasr_sbcopy
echo "installed: $(ls "$ROOT/baseline/_ref")"
