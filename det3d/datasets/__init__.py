"""Import-compatibility alias (no arithmetic): the device-side pipeline steps under the reference's module path."""
