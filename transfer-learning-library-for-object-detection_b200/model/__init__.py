"""Drop-in mirror of the reference's ``lib/model`` operator surface for the RoI /
proposal hot path.  Same dotted names, constructors and call
signatures; the work is done by libtlod_b200.so through ``tlod_b200``."""
