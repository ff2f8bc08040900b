"""One rank of the 2-GPU parity worker (tests/test_parallel_gpu.py::_worker) as a plain script, so that every rank can
run under compute-sanitizer (memcheck / racecheck) with its own log:

  for tool in memcheck racecheck; do
    for r in 0 1; do
      RANK=$r compute-sanitizer --tool $tool --log-file profiles/r02_${tool}_rank$r.log \\
          python tests/sanitize_two_gpu.py 29555 /tmp/out &
    done; wait
  done

Covers the peer-memory kernels with cross-GPU waits: k_fv_march<..., P2P> (Euler loop), k_rk_axpy / k_p2p_send_layers
(Runge-Kutta stage hand-over) and k_q1_gather<..., P2P> (interface-row halo), each checked against the CPU oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_parallel_gpu import _worker  # noqa: E402

if __name__ == "__main__":
    port, out = int(sys.argv[1]), sys.argv[2]
    os.makedirs(out, exist_ok=True)
    _worker(int(os.environ["RANK"]), 2, port, out)
    print("rank", os.environ["RANK"], "ok")
