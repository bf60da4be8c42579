"""Applications over the device cores."""
