"""Prints the driver results next to the statistical anchors committed in the reference (for DESIGN.md)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, pandas as pd
import lscpm_b200; lscpm_b200.install()
from miniplant.full_simulation import yearlong_simulation
from miniplant.locations import EINDHOVEN, NORTH_CAPE, TOWNSVILLE, PLATAFORMA_SOLAR_ALMERIA
from miniplant.solar_data import solar_data_for_place_and_time
import tempfile, time
tmp = tempfile.mkdtemp()
data = solar_data_for_place_and_time(EINDHOVEN, 40, 1800, write_csv=False)
start, end = pd.Timestamp("2020-07-15 15:00", tz=EINDHOVEN.tz), pd.Timestamp("2020-07-15 17:30", tz=EINDHOVEN.tz)
r = yearlong_simulation(40, EINDHOVEN, num_photons_per_simulation=4_000_000, time_range=(start, end), target_file=os.path.join(tmp, "w.csv"), solar_data=data, seed=1)
print("window direct ", np.round(r["simulation_direct"].values, 4), "mean", r["simulation_direct"].mean(), "| reference 0.2218 0.2191 0.2253 0.2156 0.2232 0.2230 mean 0.2213")
print("window diffuse", np.round(r["simulation_diffuse"].values, 4), "mean", r["simulation_diffuse"].mean(), "| reference 0.2087 0.2133 0.2206 0.2113 0.2215 0.2253 mean 0.2168")
for site, tilt, dye, ref in ((EINDHOVEN, 40, True, (0.1994, 0.2133)), (EINDHOVEN, 0, True, (0.1836, 0.2090)), (PLATAFORMA_SOLAR_ALMERIA, 30, True, (0.2024, 0.2139)),
                             (TOWNSVILLE, -10, True, (0.1989, 0.2132)), (NORTH_CAPE, 50, True, (0.1940, 0.2102)), (EINDHOVEN, 40, False, (0.0573, 0.0633)),
                             (NORTH_CAPE, 50, False, (0.0570, 0.0650))):
    d = solar_data_for_place_and_time(site, tilt, 1800, write_csv=False)
    t0 = time.time()
    r = yearlong_simulation(tilt, site, num_photons_per_simulation=20_000, include_dye=dye, target_file=os.path.join(tmp, "y.csv"), solar_data=d, seed=2)
    print(f"{site.name} tilt {tilt} dye={dye}: rows {len(r)} direct {r['simulation_direct'].mean():.4f} diffuse {r['simulation_diffuse'].mean():.4f} | reference {ref} | {time.time()-t0:.1f}s for {2*len(r)*20000/1e6:.0f} Mphotons")
