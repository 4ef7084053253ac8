"""Analysis plugin surface of the realignment path (reference nanopore/analyses/)."""
