"""TEST INFRASTRUCTURE ONLY.

``oracle/`` holds the CPU checker for the B200 hot path: the reference's own classes run under a
fake-omni shim (refshim.py, authoring container only), a torch-CPU restatement that travels to the GPU
box (port.py), and the script that generates tests/golden/.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import it; the product package never does.
"""
