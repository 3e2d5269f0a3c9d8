"""artis_b200 — B200-native packet propagation behind ARTIS's update_packets()."""
from .lib import ArtisB200, ArtisB200Error, library_path  # noqa: F401
from .snapshot import packets_view, read_snapshot, write_snapshot  # noqa: F401
