"""TEST INFRASTRUCTURE ONLY -- re-exports the seeded synthetic inputs (neuman_b200/synthetic.py) so the
golden generator, the oracle tests and the product benchmark all draw the same cameras / weights."""
from neuman_b200.synthetic import boost_density, camera, net_checksum, seed_nets      # noqa: F401
