"""Empty stand-in for the reference's unused plotting import."""
