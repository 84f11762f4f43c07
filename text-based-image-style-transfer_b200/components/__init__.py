"""Mirror of the reference's `components` package: only style_transfer_depth (the depth-aware / multi-plane variant of the
style-transfer loop) is built."""
