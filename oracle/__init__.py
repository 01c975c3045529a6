"""Oracle package: TEST INFRASTRUCTURE ONLY.

``oracle.orc``  - ctypes driver of the sparse C restatement (oracle/bfm_oracle.c).
``oracle.ref``  - the reference libbfm compiled unmodified by oracle/Makefile (oracle/_ref/).

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg.
The product package (bfm_b200/) never imports this.
"""
