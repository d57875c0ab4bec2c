"""Command-line helpers around the trainer (dataset preparation)."""
