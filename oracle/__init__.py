"""TEST INFRASTRUCTURE ONLY. CPU restatement of the reference hot path (liboracle.so) and, when built,
the unmodified reference itself (oracle/_ref/libcosma_ref.so). Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; cosma_b200/ never does."""
