"""Mirror of components/style_transfer_depth: the same L-BFGS loop as multi_style_transfer (single style), run once per
depth plane, with the plane split / merge on the device."""
