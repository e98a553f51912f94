"""planer_b200 (placeholder, filled in below)."""
