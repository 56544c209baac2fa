__version__ = "2.0.4+oracle"
