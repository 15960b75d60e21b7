"""oracle/ -- CPU restatement of the reference algorithm for the hot path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import anything from here, and only as the checker
or as the timed CPU baseline.  The product package (hdpgpc_b200/) never imports it and fails
loudly when its CUDA library is missing.
"""
