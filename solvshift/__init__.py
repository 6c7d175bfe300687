"""Reference package name; the implementation lives in slv_b200."""
