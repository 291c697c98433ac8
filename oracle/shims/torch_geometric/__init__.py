"""Stand-in for torch-geometric 2.5.0 (see oracle/shims/README.md). Test infrastructure only."""
__version__ = "2.5.0-shim"
