"""CPU oracle for the AudioLM hot path — TEST INFRASTRUCTURE ONLY.

Nothing in `audiolm_pytorch_b200/` may import this package.  Only `tests/`,
`__graft_entry__.smoke()` and the CPU legs of `bench.py` use it, as the checker / baseline.

Contents
  third_party.py      restatements (torch, CPU) of the two un-vendored dependencies that carry
                      hot-path arithmetic: vector-quantize-pytorch's GroupedResidualVQ (eval path)
                      and hyper-connections' HyperConnections.  PARITY UNPINNED against upstream:
                      their sources are not in /root/reference and cannot be installed offline; the
                      restatement follows the published algorithm.
  transformer.py      functional fp32 restatement of attend.py + audiolm_pytorch.py:191-1368, 1513-2137
  codec.py            functional fp32 restatement of soundstream.py:332-395, 691-709, 797-866
  ref_import.py       imports the REAL reference files from /root/reference with stub modules
                      (only works in the build container; used by make_golden.py)
  make_golden.py      runs the real reference (+ third_party.py) on seeded inputs and writes
                      tests/golden/*.pt; also asserts that transformer.py / codec.py reproduce them.
"""
