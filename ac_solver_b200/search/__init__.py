from .breadth_first import bfs, bfs_device  # noqa: F401
