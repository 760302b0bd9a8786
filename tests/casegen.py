"""Builds dsmcFoam+ case directories for the driver tests: the couette_N2-O2 tutorial re-created from the golden
fixture (mesh + cloud state shipped by the reference) with dictionaries written in the reference's keyword layout."""
import os

import numpy as np

from hystrath_b200 import capi
from hystrath_b200 import case as casew
from tests.test_oracle_golden import GOLD, couette_mesh

DSMC_PROPERTIES = """
nEquivalentParticles            %(fnum).10g;
seedNumber                      %(seed)d;

BinaryCollisionModel            LarsenBorgnakkeVariableHardSphere;

LarsenBorgnakkeVariableHardSphereCoeffs
{
    rotationalRelaxationCollisionNumber   5.0;
    inverseZvFormulation           "pre-2008";
}

collisionPartnerSelectionModel   noTimeCounter;

typeIdList                      (N2 O2);

moleculeProperties
{
    N2
    {
        mass                            46.5e-27;
        diameter                        4.17e-10;
        rotationalDegreesOfFreedom      2;
        nVibrationalModes               1;
        omega                           0.74;
        alpha                           1.0;
        characteristicVibrationalTemperature  (3371);
        dissociationTemperature         113500;
        ionisationTemperature           180798.287;
        Zref                            (52560);
        referenceTempForZref            (3371);
        charge                          0;
    }
    O2
    {
        mass                            53.12e-27;
        diameter                        4.07e-10;
        rotationalDegreesOfFreedom      2;
        nVibrationalModes               1;
        omega                           0.77;
        alpha                           1.0;
        characteristicVibrationalTemperature  (2256);   // a comment
        dissociationTemperature         59500;
        /* block
           comment */
        Zref                            (17900);
        referenceTempForZref            (2256);
        charge                          0;
    }
}
"""

CONTROL_DICT = """
application     dsmcFoam+;
nTerminalOutputs   %(nto)d;
startFrom       latestTime;
startTime       0;
stopAt          endTime;
endTime         %(end).10g;
deltaT          %(dt).10g;
writeControl    timeStep;
writeInterval   %(wi)d;
writeFormat     ascii;
writePrecision  10;
timeFormat      general;
timePrecision   10;
"""

BOUNDARIES_DICT = """
dsmcPatchBoundaries
(
    boundary
    {
        patchBoundaryProperties
        {
           patchName    upperWall;
        }
        boundaryModel   dsmcDiffuseWallPatch;
        dsmcDiffuseWallPatchProperties
        {
            temperature 		3000.0;
            velocity 			(300.0 0.0 0.0);
        }
    }
     boundary
     {
        patchBoundaryProperties
        {
           patchName    lowerWall;
        }
        boundaryModel   dsmcDiffuseWallPatch;
        dsmcDiffuseWallPatchProperties
        {
            temperature 		2000.0;
            velocity 			(0.0 0.0 0.0);
        }
     }
);

dsmcCyclicBoundaries
(
);

dsmcGeneralBoundaries
(
);
"""

FIELD_PROPERTIES = """
dsmcFields
(
     field
     {
         fieldModel             dsmcVolFields;
         timeProperties
         {
            timeOption      write;
            resetAtOutput       on;
            resetAtOutputUntilTime       0.5;
         }
         dsmcVolFieldsProperties
         {
            fieldName               O2;
            typeIds                 (O2);
            measureMeanFreePath     true;
            averagingAcrossManyRuns     false;
         }
      }
      field
      {
         fieldModel          	dsmcVolFields;
         timeProperties
         {
         	  timeOption      write;
            resetAtOutput       on;
            resetAtOutputUntilTime       0.5;
         }
         dsmcVolFieldsProperties
         {
            fieldName               N2;
            typeIds                 (N2);
            measureMeanFreePath     true;
         }
     }
      field
      {
         fieldModel             dsmcVolFields;
         timeProperties
         {
            timeOption      write;
            resetAtOutput       on;
            resetAtOutputUntilTime       0.5;
         }
         dsmcVolFieldsProperties
         {
            fieldName               mixture;
            typeIds                 (N2 O2);
            measureMeanFreePath     true;
            measureHeatFluxShearStress  true;
            measureErrors           true;
         }
     }
);
"""


def couette_case(path, n_steps=4, seed=5, nto=2, start_time="5"):
    g = np.load(GOLD)
    mesh = couette_mesh(g)
    casew.write_poly_mesh(path, mesh)
    p = capi.ParcelData(len(g["cell"]), 1, allocate=False, position=g["positions"], U=g["U"], ERot=g["ERot"], cell=g["cell"],
                        typeId=g["typeId"], vibLevel=g["vibLevel"], origId=g["origId"])
    casew.write_cloud(path, start_time, p, g["dsmcSigmaTcRMax"], mesh)
    dt = 1e-5
    casew.write_dict(os.path.join(path, "constant", "dsmcProperties"), "constant", "dsmcProperties",
                     DSMC_PROPERTIES % dict(fnum=float(g["nEquivalentParticles"]), seed=seed))
    casew.write_dict(os.path.join(path, "system", "controlDict"), "system", "controlDict",
                     CONTROL_DICT % dict(nto=nto, end=float(start_time) + n_steps * dt, dt=dt, wi=n_steps))
    casew.write_dict(os.path.join(path, "system", "boundariesDict"), "system", "boundariesDict", BOUNDARIES_DICT)
    casew.write_dict(os.path.join(path, "system", "fieldPropertiesDict"), "system", "fieldPropertiesDict", FIELD_PROPERTIES)
    return g, mesh, p


# ---- the axisymmetric tutorial (run/hyStrath/dsmcFoam+/axisymmetricFlatnosedCylinder) as a case directory ----
AXISYM_PROPERTIES = """
coordinateSystem   dsmcAxisymmetric;

nEquivalentParticles            2e7;
seedNumber                      %(seed)d;

axisymmetricProperties
{
    maxRadialWeightingFactor    1000;
}

collisionPartnerSelectionModel   		 noTimeCounter;
BinaryCollisionModel            VariableHardSphere;
VariableHardSphereCoeffs
{
    Tref        273;
}

typeIdList                      (Ar);

moleculeProperties
{
    Ar
    {
        mass                                  66.3e-27;
        diameter                              4.17e-10;
        omega                                     0.81;
        alpha                                      1.4;
    }
}
"""

AXISYM_BOUNDARIES = """
dsmcPatchBoundaries
(
     boundary
     {
         patchBoundaryProperties
         {
             patchName   cylinder;
         }
         boundaryModel   dsmcDiffuseWallPatch;
         dsmcDiffuseWallPatchProperties
         {
			      temperature 		300;
			      velocity 			(0 0 0);
         }
     }
    boundary
    {
        patchBoundaryProperties
        {
            patchName                           flow;
        }
        boundaryModel   dsmcDeletionPatch;
        dsmcDeletionPatchProperties
        {
	          allSpecies		yes;
        }
    }
);

dsmcCyclicBoundaries
(
);

dsmcGeneralBoundaries
(
    boundary
    {
        generalBoundaryProperties
        {
            patchName                           flow;
        }
        boundaryModel   dsmcFreeStreamInflowPatch;
        dsmcFreeStreamInflowPatchProperties
        {
			      typeIds						(Ar);
			      translationalTemperature           100;
		        velocity                    (1000 0 0);
		        numberDensities
		        {
		            Ar          1.0e21;
		        }
	      }
    }
);
"""

AXISYM_FIELDS = """
dsmcFields
(
	  field
    {
       fieldModel          	dsmcVolFields;
       timeProperties
       {
       	  timeOption      write;
            resetAtOutput       on;
          resetAtOutputUntilTime       8e-4;
       }
       dsmcVolFieldsProperties
       {
          fieldName                   Ar;
          typeIds                     (Ar);
          measureMeanFreePath         true;
          averagingAcrossManyRuns     false;
       }
    }
);
"""

AXISYM_INITIALISE = """
configurations
(
	configuration
  {
      type			dsmcMeshFill;
	    numberDensities
	    {
		      Ar              1.0e21;
	    };
	    translationalTemperature     	100;
	    rotationalTemperature          0;
	    vibrationalTemperature         0;
      electronicTemperature          0;
	    velocity        (1000 0 0);
	}
);
"""


def axisym_case(path, n_steps=40, seed=9, nto=10):
    """The dictionaries of the shipped tutorial (keyword layout kept) on the mesh of its blockMeshDict; dsmcInitialise+ state not yet made."""
    from hystrath_b200 import meshgen
    mesh = meshgen.axisymmetric_cylinder_mesh()
    casew.write_poly_mesh(path, mesh)
    dt = 8e-8
    casew.write_dict(os.path.join(path, "constant", "dsmcProperties"), "constant", "dsmcProperties", AXISYM_PROPERTIES % dict(seed=seed))
    casew.write_dict(os.path.join(path, "system", "controlDict"), "system", "controlDict", CONTROL_DICT % dict(nto=nto, end=n_steps * dt, dt=dt, wi=n_steps))
    casew.write_dict(os.path.join(path, "system", "boundariesDict"), "system", "boundariesDict", AXISYM_BOUNDARIES)
    casew.write_dict(os.path.join(path, "system", "fieldPropertiesDict"), "system", "fieldPropertiesDict", AXISYM_FIELDS)
    casew.write_dict(os.path.join(path, "system", "dsmcInitialiseDict"), "system", "dsmcInitialiseDict", AXISYM_INITIALISE)
    return mesh
