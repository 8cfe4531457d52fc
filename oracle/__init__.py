"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds the CPU checker for the Super4PCS hot path:

* `oracle.port`  -- ctypes binding of `oracle/liboracle_port.so`, the plain-C++ restatement
                    (`oracle/port.cc`) of the reference algorithm;
* `oracle.ref`   -- ctypes binding of `oracle/_ref/liboracle_ref.so`, the UNMODIFIED reference
                    compiled from `/root/reference` plus a thin C-ABI harness
                    (`oracle/ref_harness.cc`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package.  The product (`super4pcs_b200/`, `libs4g.so`) never does.
"""
