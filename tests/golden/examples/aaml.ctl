      seqfile = stewart.aa
     treefile = stewart.trees

      outfile = mlc
        noisy = 9
      verbose = 1
      runmode = 0


      seqtype = 2
   aaRatefile = 


        model = 0

        Mgene = 0

    fix_alpha = 1
        alpha = 0
       Malpha = 0
        ncatG = 5

        clock = 0
        getSE = 1
 RateAncestor = 1







   Small_Diff = .5e-6
    cleandata = 0
        ndata = 1
       method = 0
