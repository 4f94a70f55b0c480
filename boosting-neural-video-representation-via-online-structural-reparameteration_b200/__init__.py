"""B200-native Online-RepNeRV frame-fitting hot path.  Import through the `orepnerv` alias package at the
repository root (this directory name is not a valid Python identifier)."""
