"""Drop-in modules mirroring the reference's ``utils`` package for the registration path."""
