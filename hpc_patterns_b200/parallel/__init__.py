"""Process-group plumbing: symmetric peer memory, signal pads, topology, rank->device mapping."""
from .comm import Comm  # noqa: F401
