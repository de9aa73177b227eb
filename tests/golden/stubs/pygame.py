"""Empty stand-in for pygame (render-only dependency of the reference)."""
