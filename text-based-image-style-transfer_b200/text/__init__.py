"""Mirror of the reference's `text/` package for the pieces that sit right behind the style-transfer hot path."""
