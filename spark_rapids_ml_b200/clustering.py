"""placeholder — replaced below by the KMeans / KMeansModel surface."""
