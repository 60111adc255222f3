#!/bin/bash
# One attempt to obtain the reference's arithmetic (tensorflow 2.7 / tensorflow-addons 0.15 / h5py 3.6) on the GPU box,
# as VERDICT r01 "Next round" item 1 asks.  The box has no network and no wheelhouse entry for them, so this is expected
# to fail; the log is the committed evidence (profiles/r02_tf_install_attempt.log).  On success it would run
# oracle/tf_crosscheck.py and write TF-generated goldens.
out=${1:-gpurun_out/tf_install.log}
mkdir -p "$(dirname "$out")"
{
  echo "# $(date -u +%FT%TZ) host $(hostname)"
  python -c 'import sys; print("python", sys.version)'
  for spec in "tensorflow==2.7.0 tensorflow-addons==0.15.0 h5py==3.6.0" "tensorflow-cpu h5py" "h5py"; do
    echo "## pip install $spec"
    timeout 60 python -m pip install --timeout 5 --retries 0 --target /tmp/tf_try $spec 2>&1 | tail -8
    echo "exit: ${PIPESTATUS[0]}"
  done
  echo "## wheelhouse"
  ls /opt/wheelhouse 2>/dev/null | grep -i -E "tensorflow|h5py|keras|addons" || echo "no tensorflow / h5py / keras wheel in /opt/wheelhouse"
  echo "## import"
  PYTHONPATH=/tmp/tf_try python -c 'import tensorflow as tf; print("tensorflow", tf.__version__)' 2>&1 | tail -1
  PYTHONPATH=/tmp/tf_try python -c 'import h5py; print("h5py", h5py.__version__)' 2>&1 | tail -1
} > "$out" 2>&1
if PYTHONPATH=/tmp/tf_try python -c 'import tensorflow' 2>/dev/null; then
  PYTHONPATH=/tmp/tf_try:. python oracle/tf_crosscheck.py > gpurun_out/tf_crosscheck.log 2>&1
fi
exit 0
