"""Host-side mirror of the reference's module interface for the stage-1 hot path."""
