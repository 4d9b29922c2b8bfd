"""Mirror of ``falcon.cluster`` (reference package /root/reference/falcon/cluster)."""
from . import cluster, spectrum  # noqa: F401
