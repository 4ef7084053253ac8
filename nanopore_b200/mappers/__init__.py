"""Mapper plugin surface of the realignment path (reference nanopore/mappers/)."""
