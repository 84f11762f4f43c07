"""Drop-in mirror of the reference package `multi_style_transfer` (same module and function names):
app.py:36 imports `multi_style_transfer.run_style_transfer.run_multi_style_transfer`."""
