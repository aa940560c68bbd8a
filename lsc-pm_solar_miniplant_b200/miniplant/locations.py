"""The four sites the reference simulates (data of reference ``miniplant/locations.py:3-32``): Townsville,
Plataforma Solar de Almería, Eindhoven, North Cape — latitude / longitude in degrees, IANA time zone, altitude in m."""
try:  # real pvlib when present
    from pvlib.location import Location  # type: ignore
except Exception:  # pragma: no cover - depends on the image
    from ..compat.pvlib_lite import Location

_SITE_TABLE = (
    # constant name,             site name,                      latitude,           longitude,            time zone,            altitude
    ("TOWNSVILLE",               "Townsville",                   -19.3239872,        146.7605092,          "Australia/Brisbane", 16.3),
    ("PLATAFORMA_SOLAR_ALMERIA", "Plataforma Solar de Almería",  37.09454882268096,  -2.3586145374427008,  "Europe/Madrid",      499),
    ("EINDHOVEN",                "Eindhoven",                    51.4416,            5.6497,               "Europe/Amsterdam",   17),
    ("NORTH_CAPE",               "North Cape",                   70.976021,          25.983061,            "Europe/Stockholm",   17),
)

LOCATIONS = set()
for _constant, _name, _lat, _lon, _tz, _alt in _SITE_TABLE:
    _site = Location(latitude=_lat, longitude=_lon, tz=_tz, altitude=_alt, name=_name)
    globals()[_constant] = _site
    LOCATIONS.add(_site)
del _constant, _name, _lat, _lon, _tz, _alt, _site
