"""TEST INFRASTRUCTURE. CPU restatement ("oracle") of the NeuMan ray-marching hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (neuman_b200/) never does.
"""
