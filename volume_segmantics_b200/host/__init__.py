"""Host-side mirror of the reference's Python interface for the prediction path."""
