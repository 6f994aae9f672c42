"""galax stand-in (see ../README.md): parameter records with the reference's attribute names, no physics."""
