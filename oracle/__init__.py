"""oracle/ — CPU checkers for the mpmc++ energy hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package; the product (mpmcxx_b200/) never does and fails loudly without its CUDA library.

  oracle.port : ctypes binding of oracle/liboracle.so, our plain-C restatement (oracle.c) of the reference
                algorithm; travels in-tree, runs on the GPU box.
  oracle.ref  : ctypes binding of oracle/_ref/libmpmc_ref.so, the UNMODIFIED reference compiled from
                /root/reference/src (see Makefile); used to pin the port and to generate tests/golden/.
"""
