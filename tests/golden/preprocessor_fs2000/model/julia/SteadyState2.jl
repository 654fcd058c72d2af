function steady_state!(ys_::Vector{<: Real}, exo_::Vector{<: Real}, params::Vector{<: Real})
@inbounds begin
    ys_[14]=exp(params[3]);
    gst=1/ys_[14];
    ys_[1]=params[4];
    khst=((-gst*params[2]*(1 - params[7]) + 1)/(gst^params[1]*params[1]*params[2]))^(1/(params[1] - 1));
    xist=params[4]/(-khst*(-gst*(1 - params[7]) + 1) + (gst*khst)^params[1]);
    nust=params[4]^2*params[6]/(gst^params[1]*khst^params[1]*params[2]*(1 - params[1])*(1 - params[6]));
    ys_[9]=xist/(nust + xist);
    ys_[2]=nust + xist;
    ys_[7]=khst*ys_[9];
    ys_[10]=params[4]*params[6]*ys_[9]/((1 - params[6])*(1 - ys_[9]));
    ys_[3]=params[4]/ys_[2];
    ys_[8]=-params[4] + ys_[10] + 1;
    ys_[13]=gst^params[1]*ys_[7]^params[1]*ys_[9]^(1 - params[1]);
    ys_[6]=params[4]/params[2];
    ys_[5]=ys_[10]/ys_[9];
    ist=ys_[13] - ys_[3];
    q=1 - ys_[8];
    ys_[4]=1;
    ys_[12]=-1 + ys_[1]/ys_[14];
    ys_[11]=ys_[14] - 1;
end
end
