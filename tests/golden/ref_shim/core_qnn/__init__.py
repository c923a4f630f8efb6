"""Stand-in for the un-pinned Orkis-Research core_qnn package."""
