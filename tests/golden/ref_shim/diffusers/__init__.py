"""Stand-in for diffusers==0.27.2 (see ../README.md). Test infrastructure only."""
__version__ = "0.27.2-shim"
