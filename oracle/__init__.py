"""ORACLE / TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's sketch-guided sampling path (oracle/port.py) over a minimal
`diffusers` stand-in (oracle/diffusers_shim).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; sketch2img_b200/ never does.
"""
