"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checker for the CUDA rollout path: a plain-C restatement of the reference rules
(`loa_oracle.c`, `pan_oracle.c`) and, when /root/reference is present, the reference's own sources
compiled verbatim into the git-ignored `oracle/_ref/`.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package.  The product library
(game-ai_b200/csrc -> libgameai_gpu.so) never does and has no CPU fallback.
"""
