"""CPU oracle for the UNet++ hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this package.
The product path (pytorch_nested-unet_b200/) never does; it fails loudly when the CUDA library is missing.
"""
